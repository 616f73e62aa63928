#!/bin/sh
# Regenerates the committed golden vectors by running the UNMODIFIED reference (needs /root/reference;
# builds oracle/_ref/ref_harness with oracle/Makefile.ref first).  Seed 0xB200 = 45568.
set -e
cd "$(dirname "$0")/../.."
make -C oracle -f Makefile.ref -j8
for c in tiny_fast tiny_slow tiny_quirks tiny_full; do
  oracle/_ref/ref_harness dump "$c" "tests/golden/$c.rsgv" 45568
done
# tiny_quirks at seed 11: <s_pows, A_io> sums to a TRANSPARENT ciphertext in ring limb 0 (the seeded CRS shares its
# uniform polynomial across ciphertexts), which the reference maps to an empty zero ciphertext (seal_ring.tcc:493-504)
oracle/_ref/ref_harness dump tiny_quirks tests/golden/tiny_transp.rsgv 11
