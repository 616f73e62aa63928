#!/bin/sh
# Regenerates the committed golden vectors by running the UNMODIFIED reference (needs /root/reference;
# builds oracle/_ref/ref_harness with oracle/Makefile.ref first).  Seed 0xB200 = 45568.
set -e
cd "$(dirname "$0")/../.."
make -C oracle -f Makefile.ref -j8
for c in tiny_fast tiny_slow tiny_quirks tiny_full; do
  oracle/_ref/ref_harness dump "$c" "tests/golden/$c.rsgv" 45568
done
