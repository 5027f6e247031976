
// ---- appended by aeonflux_b200's oracle/_ref_recipe/run.sh -- NOT part of the reference -------------------------------------
// Golden-vector dumper.  Runs the reference's own public flow (SystemParameters::generate -> Issuer::new ->
// CredentialRequestConstructor -> Issuer::issue -> CredentialIssuance::verify -> hide_attribute -> AnonymousCredential::show ->
// Issuer::verify) with a deterministic byte-stream rng and writes, per case, the issuer's to_bytes() encodings, every
// presentation / issuance flattened to 32-byte words in the aeonflux_b200 layout (include/aeonflux_b200.h), the reference's
// verdicts, and the verdicts on deliberately corrupted copies.  Output: $B200_VECTORS_OUT/ref_<case>.json (default ".").
//
// What the stream does NOT determine: the blindings zkp's prove_compact draws (TranscriptRng finalised with thread_rng), hence
// the challenge / response words.  Everything else (parameters, keys, attributes, t, U, V, z, every commitment point, the
// ciphertexts) is a function of the stream and is replayed byte for byte by tests/test_reference_vectors.py.
//
//     cargo test --release b200_vectors -- --nocapture
#[cfg(test)]
mod b200_vectors {
    use std::fs::File;
    use std::io::Write;
    use std::string::String;
    use std::vec::Vec;

    use curve25519_dalek::constants::RISTRETTO_BASEPOINT_POINT;
    use curve25519_dalek::ristretto::RistrettoPoint;
    use curve25519_dalek::scalar::Scalar;
    use curve25519_dalek::traits::Identity;

    use rand_core::CryptoRng;
    use rand_core::RngCore;

    use sha2::Digest;
    use sha2::Sha512;

    use crate::amacs::Attribute;
    use crate::amacs::EncryptedAttribute;
    use crate::credential::AnonymousCredential;
    use crate::issuer::CredentialIssuance;
    use crate::issuer::Issuer;
    use crate::parameters::SystemParameters;
    use crate::symmetric::Keypair as SymmetricKeypair;
    use crate::user::CredentialRequestConstructor;

    use super::ProofOfValidCredential;

    /// Deterministic rng: block i of the stream is SHA-512(seed || le64(i)); every request takes the next bytes.
    struct StreamRng {
        seed: Vec<u8>,
        counter: u64,
        block: Vec<u8>,
        position: usize,
        taken: u64,
    }

    impl StreamRng {
        fn new(seed: &[u8]) -> StreamRng {
            StreamRng { seed: seed.to_vec(), counter: 0, block: Vec::new(), position: 0, taken: 0 }
        }

        fn refill(&mut self) {
            let mut h = Sha512::new();

            h.input(&self.seed[..]);
            h.input(&self.counter.to_le_bytes()[..]);
            self.block = h.result().to_vec();
            self.position = 0;
            self.counter += 1;
        }
    }

    impl RngCore for StreamRng {
        fn next_u32(&mut self) -> u32 {
            let mut b = [0u8; 4];
            self.fill_bytes(&mut b);
            u32::from_le_bytes(b)
        }

        fn next_u64(&mut self) -> u64 {
            let mut b = [0u8; 8];
            self.fill_bytes(&mut b);
            u64::from_le_bytes(b)
        }

        fn fill_bytes(&mut self, dest: &mut [u8]) {
            for d in dest.iter_mut() {
                if self.position == self.block.len() {
                    self.refill();
                }
                *d = self.block[self.position];
                self.position += 1;
                self.taken += 1;
            }
        }

        fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> {
            self.fill_bytes(dest);
            Ok(())
        }
    }

    impl CryptoRng for StreamRng {}

    #[derive(Clone, Copy, PartialEq)]
    enum Req {
        /// Scalar::random(rng) appended as a revealed scalar.
        PS,
        /// RistrettoPoint::random(rng) appended as a revealed point.
        PP,
        /// 30 rng bytes appended as a plaintext.
        EP,
        /// The all-zero 30-byte plaintext (its M1 is the identity, issuance.rs:271-295); no rng bytes.
        EPZ,
    }

    fn hex(bytes: &[u8]) -> String {
        let mut s = String::new();

        for b in bytes.iter() {
            s.push_str(&format!("{:02x}", b));
        }
        s
    }

    fn json_words(words: &Vec<[u8; 32]>) -> String {
        let mut s = String::from("[");

        for (i, w) in words.iter().enumerate() {
            if i > 0 {
                s.push_str(", ");
            }
            s.push_str(&format!("\"{}\"", hex(&w[..])));
        }
        s.push_str("]");
        s
    }

    fn json_u8s(v: &Vec<u8>) -> String {
        let mut s = String::from("[");

        for (i, x) in v.iter().enumerate() {
            if i > 0 {
                s.push_str(", ");
            }
            s.push_str(&format!("{}", x));
        }
        s.push_str("]");
        s
    }

    /// kinds (0 revealed scalar, 1 hidden scalar, 2 revealed point, 3 hidden point) and the flat words:
    /// challenge, responses, C_x_0, C_x_1, C_V, C_y[n], revealed values in index order, then 14 words per proof of encryption.
    fn presentation_words(p: &ProofOfValidCredential) -> (Vec<u8>, Vec<[u8; 32]>) {
        let mut kinds: Vec<u8> = Vec::new();
        let mut w: Vec<[u8; 32]> = Vec::new();

        w.push(p.proof.challenge.to_bytes());
        for r in p.proof.responses.iter() {
            w.push(r.to_bytes());
        }
        w.push(p.C_x_0.compress().to_bytes());
        w.push(p.C_x_1.compress().to_bytes());
        w.push(p.C_V.compress().to_bytes());
        for c in p.C_y.iter() {
            w.push(c.compress().to_bytes());
        }
        for a in p.encrypted_attributes.iter() {
            match a {
                EncryptedAttribute::PublicScalar(m) => { kinds.push(0); w.push(m.to_bytes()); },
                EncryptedAttribute::SecretScalar    => { kinds.push(1); },
                EncryptedAttribute::PublicPoint(M)  => { kinds.push(2); w.push(M.compress().to_bytes()); },
                EncryptedAttribute::SecretPoint     => { kinds.push(3); },
            }
        }
        for (_index, proof_of_encryption) in p.proofs_of_encryption.iter() {
            for x in proof_of_encryption.b200_words().iter() {
                w.push(*x);
            }
        }
        (kinds, w)
    }

    /// kinds (0 scalar attribute, 2 point attribute) and the flat words: attribute[n], t, U, V, challenge, responses[n + 5].
    fn issuance_words(issuance: &CredentialIssuance) -> (Vec<u8>, Vec<[u8; 32]>) {
        let mut kinds: Vec<u8> = Vec::new();
        let mut w: Vec<[u8; 32]> = Vec::new();

        for a in issuance.credential.attributes.iter() {
            match a {
                Attribute::PublicScalar(m) => { kinds.push(0); w.push(m.to_bytes()); },
                Attribute::SecretScalar(m) => { kinds.push(0); w.push(m.to_bytes()); },
                Attribute::PublicPoint(M)  => { kinds.push(2); w.push(M.compress().to_bytes()); },
                Attribute::EitherPoint(p)  => { kinds.push(2); w.push(p.M1.compress().to_bytes()); },
                Attribute::SecretPoint(p)  => { kinds.push(2); w.push(p.M1.compress().to_bytes()); },
            }
        }
        w.push(issuance.credential.amac.t.to_bytes());
        w.push(issuance.credential.amac.U.compress().to_bytes());
        w.push(issuance.credential.amac.V.compress().to_bytes());
        for x in issuance.proof.b200_words().iter() {
            w.push(*x);
        }
        (kinds, w)
    }

    fn verdict_of<T>(r: Result<T, crate::errors::CredentialError>) -> u8 {
        match r {
            Ok(_) => 0,
            Err(_) => 1,
        }
    }

    fn corrupted_entry(class: &str, p: &ProofOfValidCredential, issuer: &Issuer) -> String {
        let (_kinds, words) = presentation_words(p);
        let verdict = verdict_of(p.verify(issuer));

        format!("{{\"class\": \"{}\", \"verdict\": {}, \"words\": {}}}", class, verdict, json_words(&words))
    }

    /// Every corruption class the Rust types can hold, applied to `p` in place, verified, recorded and undone.
    fn presentation_corruptions(p: &mut ProofOfValidCredential, issuer: &Issuer) -> Vec<String> {
        let mut out: Vec<String> = Vec::new();
        let one = Scalar::one();
        let B = RISTRETTO_BASEPOINT_POINT;

        p.proof.responses[0] = p.proof.responses[0] + one;
        out.push(corrupted_entry("response+1", p, issuer));
        p.proof.responses[0] = p.proof.responses[0] - one;

        p.proof.challenge = p.proof.challenge + one;
        out.push(corrupted_entry("challenge+1", p, issuer));
        p.proof.challenge = p.proof.challenge - one;

        p.C_x_0 = p.C_x_0 + B;
        out.push(corrupted_entry("C_x_0+B", p, issuer));
        p.C_x_0 = p.C_x_0 - B;

        p.C_x_1 = p.C_x_1 + B;
        out.push(corrupted_entry("C_x_1+B", p, issuer));
        p.C_x_1 = p.C_x_1 - B;

        p.C_V = p.C_V + B;
        out.push(corrupted_entry("C_V+B", p, issuer));
        p.C_V = p.C_V - B;

        let saved = p.C_y[0];
        p.C_y[0] = RistrettoPoint::identity();
        out.push(corrupted_entry("identity_point", p, issuer));
        p.C_y[0] = saved + B;
        out.push(corrupted_entry("C_y_0+B", p, issuer));
        p.C_y[0] = saved;

        for i in 0..p.encrypted_attributes.len() {
            let original = p.encrypted_attributes[i].clone();

            match original {
                EncryptedAttribute::PublicScalar(m) => {
                    p.encrypted_attributes[i] = EncryptedAttribute::PublicScalar(m + one);
                    out.push(corrupted_entry("revealed_scalar", p, issuer));
                },
                EncryptedAttribute::PublicPoint(M) => {
                    p.encrypted_attributes[i] = EncryptedAttribute::PublicPoint(M + B);
                    out.push(corrupted_entry("revealed_point", p, issuer));
                },
                _ => continue,
            }
            p.encrypted_attributes[i] = original;
        }

        if p.proofs_of_encryption.len() > 0 {
            let last = p.proofs_of_encryption.len() - 1;

            p.proofs_of_encryption[0].1.b200_corrupt(0, false);
            out.push(corrupted_entry("enc_E2+B", p, issuer));
            p.proofs_of_encryption[0].1.b200_corrupt(0, true);

            p.proofs_of_encryption[last].1.b200_corrupt(1, false);
            out.push(corrupted_entry("enc_response+1", p, issuer));
            p.proofs_of_encryption[last].1.b200_corrupt(1, true);
        }

        // everything was undone: the presentation verifies as it did before
        out
    }

    fn issuance_corruptions(issuance: &mut CredentialIssuance, issuer: &Issuer) -> Vec<String> {
        let mut out: Vec<String> = Vec::new();
        let sp = &issuer.system_parameters;
        let ip = &issuer.issuer_parameters;
        let n = issuance.credential.attributes.len();

        fn entry(class: &str, issuance: &CredentialIssuance, sp: &SystemParameters, ip: &crate::parameters::IssuerParameters) -> String {
            let (_kinds, words) = issuance_words(issuance);
            let verdict = verdict_of(issuance.proof.verify(sp, ip, &issuance.credential));

            format!("{{\"class\": \"{}\", \"verdict\": {}, \"words\": {}}}", class, verdict, json_words(&words))
        }

        issuance.proof.b200_corrupt(0, 0, false);
        out.push(entry("response_w+1", issuance, sp, ip));
        issuance.proof.b200_corrupt(0, 0, true);

        issuance.proof.b200_corrupt(0, n + 4, false);
        out.push(entry("response_one+1", issuance, sp, ip));
        issuance.proof.b200_corrupt(0, n + 4, true);

        issuance.proof.b200_corrupt(1, 0, false);
        out.push(entry("challenge+1", issuance, sp, ip));
        issuance.proof.b200_corrupt(1, 0, true);

        issuance.credential.amac.V = issuance.credential.amac.V + RISTRETTO_BASEPOINT_POINT;
        out.push(entry("V+B", issuance, sp, ip));
        issuance.credential.amac.V = issuance.credential.amac.V - RISTRETTO_BASEPOINT_POINT;

        issuance.credential.amac.t = issuance.credential.amac.t + Scalar::one();
        out.push(entry("t+1", issuance, sp, ip));
        issuance.credential.amac.t = issuance.credential.amac.t - Scalar::one();

        out
    }

    fn dump_case(name: &str, n: u32, request: &[Req], hide: &[usize], items: usize) {
        let mut seed: Vec<u8> = b"aeonflux-b200/reference-vectors/".to_vec();

        seed.extend_from_slice(name.as_bytes());

        let mut rng = StreamRng::new(&seed[..]);
        let system_parameters = SystemParameters::generate(&mut rng, n).unwrap();
        let issuer = Issuer::new(&system_parameters, &mut rng);
        let mut issuer_pub: Vec<u8> = Vec::new();

        issuer_pub.extend_from_slice(&issuer.issuer_parameters.C_W.compress().to_bytes()[..]);
        issuer_pub.extend_from_slice(&issuer.issuer_parameters.I.compress().to_bytes()[..]);

        let mut entries: Vec<String> = Vec::new();

        for item in 0..items {
            let stream_start = rng.taken;
            let mut constructor = CredentialRequestConstructor::new(&system_parameters);

            for r in request.iter() {
                match r {
                    Req::PS => constructor.append_revealed_scalar(Scalar::random(&mut rng)),
                    Req::PP => constructor.append_revealed_point(RistrettoPoint::random(&mut rng)),
                    Req::EP => {
                        let mut message = [0u8; 30];

                        rng.fill_bytes(&mut message);
                        let _plaintexts = constructor.append_plaintext(&message.to_vec());
                    },
                    Req::EPZ => {
                        let _plaintexts = constructor.append_plaintext(&vec![0u8; 30]);
                    },
                }
            }

            let mut issuance = issuer.issue(constructor.finish(), &mut rng).unwrap();
            let (issuance_kinds, issuance_w) = issuance_words(&issuance);
            let issuance_verdict = verdict_of(issuance.proof.verify(&system_parameters, &issuer.issuer_parameters, &issuance.credential));
            let issuance_corrupted = issuance_corruptions(&mut issuance, &issuer);
            let mut credential: AnonymousCredential = issuance.credential.clone();
            let (keypair, _master_secret) = SymmetricKeypair::generate(&system_parameters, &mut rng);

            for index in hide.iter() {
                credential.hide_attribute(*index).unwrap();
            }

            let mut presentation = credential.show(&system_parameters, &issuer.issuer_parameters, Some(&keypair), &mut rng).unwrap();
            let (kinds, words) = presentation_words(&presentation);
            let verdict = verdict_of(issuer.verify(&presentation));
            let corrupted = presentation_corruptions(&mut presentation, &issuer);

            assert!(verdict == verdict_of(issuer.verify(&presentation)));

            entries.push(format!(
                "{{\"item\": {}, \"stream_start\": {}, \"stream_end\": {}, \"kinds\": {}, \"words\": {}, \"verdict\": {}, \"corrupted\": [{}],\n  \
                 \"issuance_kinds\": {}, \"issuance_words\": {}, \"issuance_verdict\": {}, \"issuance_corrupted\": [{}]}}",
                item, stream_start, rng.taken, json_u8s(&kinds), json_words(&words), verdict, corrupted.join(",\n   "),
                json_u8s(&issuance_kinds), json_words(&issuance_w), issuance_verdict, issuance_corrupted.join(",\n   ")));
        }

        let request_names: Vec<String> = request.iter().map(|r| String::from(match r {
            Req::PS => "\"PS\"", Req::PP => "\"PP\"", Req::EP => "\"EP\"", Req::EPZ => "\"EPZ\"",
        })).collect();
        let hide_u8: Vec<u8> = hide.iter().map(|i| *i as u8).collect();
        let json = format!(
            "{{\"source\": \"isislovecruft/aeonflux 0.2.0 (the reference crate itself), b200_vectors test module\",\n \
             \"name\": \"{}\", \"n\": {}, \"seed\": \"{}\", \"request\": [{}], \"hide\": {},\n \
             \"sysparams\": \"{}\",\n \"issuer_pub\": \"{}\",\n \"secret\": \"{}\",\n \"items\": [\n  {}\n ]}}\n",
            name, n, hex(&seed[..]), request_names.join(", "), json_u8s(&hide_u8),
            hex(&system_parameters.to_bytes()[..]), hex(&issuer_pub[..]), hex(&issuer.amacs_key.to_bytes()[..]),
            entries.join(",\n  "));
        let directory = match std::env::var("B200_VECTORS_OUT") {
            Ok(d) => d,
            Err(_) => String::from("."),
        };
        let path = format!("{}/ref_{}.json", directory, name);
        let mut file = File::create(&path).unwrap();

        file.write_all(json.as_bytes()).unwrap();
        println!("b200_vectors: wrote {}", path);
    }

    #[test]
    fn b200_vectors() {
        use self::Req::*;

        // BASELINE configs[0..1]: the README flow's shape (README.md:44-117), attributes 0 and 3 hidden
        dump_case("readme4", 4, &[PS, PS, PP, EP], &[0, 3], 4);
        // BASELINE configs[3]: 16 attributes, 8 hidden plaintexts last
        dump_case("s16", 16, &[PS, PS, PS, PS, PS, PS, PP, PP, EP, EP, EP, EP, EP, EP, EP, EP], &[0, 1, 8, 9, 10, 11, 12, 13, 14, 15], 2);
        // shapes of the reference's own tests (presentation.rs:460-638)
        dump_case("revealed10", 10, &[PP, PP, PS, PS, PP, PS, PP, PS, PS, PP], &[], 1);
        dump_case("plain10_hidden_scalar", 10, &[EP, PP, PS, PS, PP, PS, PP, PS, PS, PP], &[2], 1);
        dump_case("plain1_hidden", 1, &[EP], &[0], 1);
        dump_case("scalar1", 1, &[PS], &[], 1);
        // compacted-index quirk (presentation.rs:427-433): a hidden plaintext first always fails, in the middle it passes
        dump_case("quirk_sp_first", 3, &[EP, PS, PS], &[0], 1);
        dump_case("quirk_sp_middle", 3, &[PS, EP, PS], &[1], 1);
        // issuance.rs:271-295: an identity-valued message fails CredentialIssuance::verify
        dump_case("identity_plaintext", 6, &[EPZ, PS, PS, PP, PP, PS], &[], 1);
    }
}
