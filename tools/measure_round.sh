#!/bin/bash
# The evidence set of a round, produced by ONE gpurun call (one box):  tools/measure_round.sh <tag>
# Outputs land in gpurun_out/<tag>_*; the ones to be judged are copied into profiles/ by hand.
tag=${1:-rX}
o=gpurun_out
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -2 $o/${tag}_bench.err
python bench.py --impl reference --steps 6 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
python bench.py --mode stream --frames 300 > $o/${tag}_bench_stream.json 2>> $o/${tag}_bench.err
# launch list of the bench's own steady-state steps (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
    --log-file $o/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --prefill 13 --only-device --no-cpu-baseline \
    --cuda-profiler > $o/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $o/${tag}_launches.csv 4 > $o/${tag}_launch_summary.txt
# full-section capture of the two dominant kernels, one launch each
ncu --set full --clock-control none --import-source on -k 'regex:flash_d256_tcgen05_kernel|gemm2_bf16_tcgen05_2cta_kernel' -c 6 \
    -o $o/${tag}_full -f python tools/kernel_probe.py --once --only "e4.fc1,ma.ff1,flash cross B16 N28736" \
    > $o/${tag}_ncu_full.log 2>&1
tail -2 $o/${tag}_ncu_full.log
head -30 $o/${tag}_launch_summary.txt
cut -c1-1500 $o/${tag}_bench.json
cut -c1-400 $o/${tag}_bench_stream.json
cut -c1-600 $o/${tag}_bench_reference.json
# warm per-seam and per-op times of the same configuration (CUDA events; graphs on for the seams, off for the op trace)
python tools/seam_times.py > $o/${tag}_seam_times.json 2>/dev/null
python tools/op_trace.py --prefill 8 --steps 3 --out $o/${tag}_op_trace.txt > /dev/null 2>&1
tail -1 $o/${tag}_seam_times.json | cut -c1-700
