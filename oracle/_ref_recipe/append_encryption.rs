
// ---- appended by aeonflux_b200's oracle/_ref_recipe/run.sh -- NOT part of the reference -------------------------------------
// Test-only accessors for the golden-vector dumper (`mod b200_vectors` appended to presentation.rs): the fields of
// ProofOfEncryption are private to this module.
#[cfg(test)]
impl ProofOfEncryption {
    /// The 14 words of the flat layout: challenge, responses[6], pk, E1, E2, C_y_1, C_y_2, C_y_3, C_y_2'.
    pub(crate) fn b200_words(&self) -> std::vec::Vec<[u8; 32]> {
        let mut w: std::vec::Vec<[u8; 32]> = std::vec::Vec::new();

        w.push(self.proof.challenge.to_bytes());
        for r in self.proof.responses.iter() {
            w.push(r.to_bytes());
        }
        w.push(self.public_key.pk.compress().to_bytes());
        w.push(self.ciphertext.E1.compress().to_bytes());
        w.push(self.ciphertext.E2.compress().to_bytes());
        w.push(self.C_y_1.compress().to_bytes());
        w.push(self.C_y_2.compress().to_bytes());
        w.push(self.C_y_3.compress().to_bytes());
        w.push(self.C_y_2_prime.compress().to_bytes());
        w
    }

    pub(crate) fn b200_index(&self) -> u16 {
        self.index
    }

    /// class 0: E2 +/- B; class 1: responses[0] +/- 1.  `undo` reverses a previous call.
    pub(crate) fn b200_corrupt(&mut self, class: u8, undo: bool) {
        use curve25519_dalek::constants::RISTRETTO_BASEPOINT_POINT;

        match (class, undo) {
            (0, false) => self.ciphertext.E2 = self.ciphertext.E2 + RISTRETTO_BASEPOINT_POINT,
            (0, true)  => self.ciphertext.E2 = self.ciphertext.E2 - RISTRETTO_BASEPOINT_POINT,
            (_, false) => self.proof.responses[0] = self.proof.responses[0] + Scalar::one(),
            (_, true)  => self.proof.responses[0] = self.proof.responses[0] - Scalar::one(),
        }
    }
}
