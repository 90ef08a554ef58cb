#!/bin/bash
# frame source on the device: simulation speed of the C driver with both sources
mkdir -p /tmp/mc/data && cd /tmp/mc
M=$GRAFT_REPO_ROOT/oracle/_ref/matrices
X=$GRAFT_REPO_ROOT/ems-decoder-of-nb-ldpc-codes_b200/nbldpc_mc
run() { echo "== $*"; timeout 600 $X "$@" 2>&1 | tr '\r' '\n' | grep -E "FER=|frame source" | tail -2; }
run 6000 10 $M/AD_64800_R12_GF256 2.6 20 0.3 25 2368 0 0 1
run 600 10 $M/AD_64800_R12_GF256 2.6 20 0.3 25 600 0 0 0
run 6000 10 $M/AD_64800_R12_GF256 3.0 20 0.3 25 2368 0 0 1
run 6000 10 $M/Ahmed_64800_R34_GF16 3.2 16 0.3 25 2368 0 0 1
run 300 10 $M/Ahmed_64800_R34_GF16 3.2 16 0.3 25 300 0 0 0
run 2000000 10 $M/N96_K48_GF64 4.5 20 0.3 25 65536 0 0 1
run 100000 10 $M/N96_K48_GF64 4.5 20 0.3 25 65536 0 0 0
run 200000 10 $M/Mat24_N480_M240 2.5 16 0.3 25 16384 0 0 1
