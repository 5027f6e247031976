#!/bin/bash
# Produces tests/golden/ref_*.json with the REFERENCE CRATE ITSELF (isislovecruft/aeonflux 0.2.0).
#
#   oracle/_ref_recipe/run.sh /path/to/aeonflux-checkout        (needs rustup + network access to crates.io; neither exists
#                                                                in the build image, which is why the output is not committed yet)
#
# What it does -- nothing is copied from the reference into this repository:
#   1. copies the checkout to a scratch directory (the reference tree is never modified);
#   2. APPENDS three test-only blocks to the copy: accessors for the private proof fields (append_encryption.rs, append_issuance.rs)
#      and the dumper `#[cfg(test)] mod b200_vectors` (append_presentation.rs) -- appending needs no context lines of the reference;
#   3. drops the criterion dev-dependency and the stale [[bench]] target (benches/ does not compile against the current API);
#   4. pins the toolchain (the crate needs `#![feature(try_trait)]`, src/lib.rs:25: a nightly from before 2021-05) and the
#      dependency versions (pins.txt; a Cargo.lock committed beside this script takes precedence);
#   5. runs `cargo test --release b200_vectors`, which writes the vectors; copies the Cargo.lock it used back here.
# Then:  python -m pytest tests/test_reference_vectors.py      (and -m gpu on a B200)  -- and commit tests/golden/ref_*.json.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
SRC="${1:?usage: run.sh /path/to/aeonflux-checkout}"
TOOLCHAIN="$(cat "$HERE/rust-toolchain")"
WORK="$(mktemp -d)"
cp -r "$SRC"/. "$WORK"/
cd "$WORK"
rm -rf .git target benches
cat "$HERE/append_encryption.rs"   >> src/nizk/encryption.rs
cat "$HERE/append_issuance.rs"     >> src/nizk/issuance.rs
cat "$HERE/append_presentation.rs" >> src/nizk/presentation.rs
# the stale bench target and its dev-dependency (criterion pulls ~100 crates that no longer build on a 2021 nightly)
python3 - <<'PY'
import re
s = open("Cargo.toml").read()
s = re.sub(r"\[\[bench\]\]\nname = \"aeonflux_benchmarks\"\nharness = false\n", "", s)
s = re.sub(r"criterion = \{[^}]*\}\n", "", s)
open("Cargo.toml", "w").write(s)
PY
cp "$HERE/rust-toolchain" rust-toolchain
rustup toolchain install "$TOOLCHAIN" --profile minimal
if [ -f "$HERE/Cargo.lock" ]; then
    cp "$HERE/Cargo.lock" Cargo.lock
else
    cargo generate-lockfile
    while read -r crate version; do
        case "$crate" in ""|\#*) continue;; esac
        cargo update -p "$crate" --precise "$version" || echo "run.sh: could not pin $crate to $version (continuing)"
    done < "$HERE/pins.txt"
fi
mkdir -p "$REPO/tests/golden"
B200_VECTORS_OUT="$REPO/tests/golden" cargo test --release b200_vectors -- --nocapture
cp Cargo.lock "$HERE/Cargo.lock"
ls -l "$REPO"/tests/golden/ref_*.json
echo "now run: python -m pytest tests/test_reference_vectors.py -q   (the skips must be gone)"
