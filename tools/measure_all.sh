#!/bin/bash
# Measurement campaign of a build (run under gpurun from the repo root); everything lands in gpurun_out/<tag>_*.
# usage: tools/measure_all.sh <tag> [tests] [bench] [ncu] [sanitizer]
tag=$1; shift
out=gpurun_out
for what in "$@"; do
case $what in
tests)
  (timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_gputest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gputest.log); tail -3 $out/${tag}_gputest.log
  (timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log); tail -3 $out/${tag}_smoke.log | cut -c1-700
  ;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err; echo "cfg5 rc=$?"
  timeout 600 python bench.py --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_b8.json 2> $out/${tag}_bench_b8.err; echo "b8 rc=$?"
  timeout 600 python bench.py --batch 8 --steps 400 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_b8_sustained.json 2> $out/${tag}_bench_b8_sustained.err; echo "b8 sustained rc=$?"
  timeout 600 python bench.py --workload sif --batch 32 --steps 5 --warmup 3 > $out/${tag}_bench_sif_b32.json 2> $out/${tag}_bench_sif_b32.err; echo "sif rc=$?"
  timeout 600 python bench.py --workload enc --batch 64 --steps 5 --warmup 3 > $out/${tag}_bench_enc_b64.json 2> $out/${tag}_bench_enc_b64.err; echo "enc rc=$?"
  timeout 600 python bench.py --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --precision dec1_y1 > $out/${tag}_bench_b8_dec1_y1.json 2> /dev/null; echo "y1 rc=$?"
  timeout 600 python bench.py --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --precision exact > $out/${tag}_bench_b8_exact.json 2> /dev/null; echo "exact rc=$?"
  timeout 300 python bench.py --workload codec --steps 5 --warmup 3 > $out/${tag}_bench_codec.json 2> /dev/null; echo "codec rc=$?"
  timeout 300 python bench.py --workload roundtrip --steps 3 --warmup 3 > $out/${tag}_bench_roundtrip.json 2> /dev/null; echo "roundtrip rc=$?"
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> /dev/null; echo "reference rc=$?"
  ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches.csv python tools/ncu_targets.py --what step > $out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $out/${tag}_full python tools/ncu_targets.py --what sinet,probclass,quant,sif,trunk,enc > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ;;
sanitizer)
  timeout 480 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_conv_ws.py tests/test_gpu_codec.py tests/test_gpu_sifinder_edge.py -q -x -k "shape0 or shape2 or all_nan or identical or small or random" > $out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $out/${tag}_memcheck.log
  timeout 480 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_conv_ws.py tests/test_gpu_codec.py -q -x -k "shape0 or shape2 or small or random" > $out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $out/${tag}_racecheck.log
  ;;
esac
done
