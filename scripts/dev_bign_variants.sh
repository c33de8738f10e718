#!/bin/bash
# Dev: large-n throughput of every variant library under altro_mpc_icra2021_b200/csrc/var/ (ALTRO_B200_LIB override)
out=gpurun_out/${1:-variants}.md
echo "| lib | n | m | line |" > $out
for lib in altro_mpc_icra2021_b200/csrc/libaltro_b200.so altro_mpc_icra2021_b200/csrc/var/*.so; do
  for nm in "64 16" "128 32" "200 25"; do
    set -- $nm
    r=$(ALTRO_B200_LIB=$PWD/$lib NN=$1 MM=$2 B=296 K=4 timeout 300 python scripts/dev_bign.py 2>&1 | tail -1 | sed 's/.*} //')
    echo "| $(basename $lib) | $1 | $2 | $r |" >> $out
  done
done
