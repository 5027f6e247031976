
// ---- appended by aeonflux_b200's oracle/_ref_recipe/run.sh -- NOT part of the reference -------------------------------------
// Test-only accessors for the golden-vector dumper: the CompactProof inside ProofOfIssuance is private to this module.
#[cfg(test)]
impl ProofOfIssuance {
    /// challenge, responses[n + 5]
    pub(crate) fn b200_words(&self) -> std::vec::Vec<[u8; 32]> {
        let mut w: std::vec::Vec<[u8; 32]> = std::vec::Vec::new();

        w.push(self.0.challenge.to_bytes());
        for r in self.0.responses.iter() {
            w.push(r.to_bytes());
        }
        w
    }

    /// class 0: responses[k] +/- 1; class 1: challenge +/- 1.
    pub(crate) fn b200_corrupt(&mut self, class: u8, k: usize, undo: bool) {
        match (class, undo) {
            (0, false) => self.0.responses[k] = self.0.responses[k] + Scalar::one(),
            (0, true)  => self.0.responses[k] = self.0.responses[k] - Scalar::one(),
            (_, false) => self.0.challenge = self.0.challenge + Scalar::one(),
            (_, true)  => self.0.challenge = self.0.challenge - Scalar::one(),
        }
    }
}
