#!/bin/bash
# Runs on the GPU box (under gpurun, one GPU): launch list + one full capture per dominant kernel of bench.py.
# usage: bash profiles/capture.sh <tag>      outputs land in gpurun_out/
set -u
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_${TAG}_plain.json 2> $O/bench_${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${TAG}_bench_launches_ncu.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_${TAG}_launches.log 2>&1
cap() {  # name, kernel regex, skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/prof_${TAG}_$1 -f \
      python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_${TAG}_$1.log 2>&1
}
cap sha3_short sha3_short 5
cap sponge_kmac 'sponge_kernel<' 2
cap sponge_chain sponge_chain_kernel 1
cap sponge_ae 'sponge_kernel2' 1
cap sponge_tiered sponge_tiered 0
cap fixed_base fixed_base_kernel 1
cap var_base var_base_kernel 1
python profiles/summarise_ncu.py $O/prof_${TAG}_*.ncu-rep > $O/${TAG}_ncu_summaries.txt
# gpurun brings back at most 64 MiB of gpurun_out/: keep the summaries, drop the reports
rm -f $O/prof_${TAG}_*.ncu-rep
ls -la $O | tail -20
