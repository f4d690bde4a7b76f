#!/bin/bash
FA=/dev/shm/c3s.fa
tools/_build/gen_reads -n 5000 -L 10000 -G 460000 -m ont -s 20240604 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000"
for d in 2 2 2 0 0 1 1; do echo depth $d; ZMO_DEPTH=$d $W $ARGS 2>&1 | tail -1 | cut -c60-175; done
