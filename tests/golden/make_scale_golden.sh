#!/bin/bash
# Regenerates tests/golden/scale_digests.json: the UNMODIFIED reference binary (oracle/_ref/wtzmo -t 1, ~8 min of CPU) on one
# query shard of a 200,000-read x 10 kb set (2.09 Gbp, 20 Mb genome).  Run where /root/reference exists.
set -e
HERE=$(cd "$(dirname "$0")" && pwd); REPO=$(cd "$HERE/../.." && pwd)
make -C "$REPO/oracle" ref >/dev/null
T=$(mktemp -d)
"$REPO/tools/_build/gen_reads" -n 200000 -L 10000 -G 20000000 -m pacbio -s 20240605 -o "$T/big.fa"
"$REPO/oracle/_ref/wtzmo" -t 1 -i "$T/big.fa" -fo "$T/ref.ovl" -k 16 -s 200 -m 0.6 -P 400 -p 0 2>/dev/null
printf '{\n "big200k_P400_p0": {"gen": ["-n", "200000", "-L", "10000", "-G", "20000000", "-m", "pacbio", "-s", "20240605"], "args": ["-k", "16", "-s", "200", "-m", "0.6", "-P", "400", "-p", "0"], "md5": "%s", "lines": %d, "contained_md5": "%s"}\n}\n' \
  "$(md5sum < "$T/ref.ovl" | cut -d' ' -f1)" "$(wc -l < "$T/ref.ovl")" "$(md5sum < "$T/ref.ovl.contained" | cut -d' ' -f1)" > "$HERE/scale_digests.json"
cat "$HERE/scale_digests.json"; rm -rf "$T"
