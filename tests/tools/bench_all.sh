# round-end measurement set on one B200: every workload through bench.py (GPU arm), the reference arm of the headline
# workload, and the ncu launch list of the headline command (run plainly first, as B200_PROFILING.md asks)
set -x
for wl in teapot1080 dragon1080_primary analytic1080 dragon16_8k; do
  python bench.py --steps 20 --workload $wl > gpurun_out/r02_bench_${wl}_1gpu.json 2> gpurun_out/r02_bench_${wl}_1gpu.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_dragon4k_reference_arm.json 2> gpurun_out/r02_bench_dragon4k_reference_arm.err
python bench.py --steps 20 > gpurun_out/r02_bench_dragon4k_1gpu.json 2> gpurun_out/r02_bench_dragon4k_1gpu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-render > gpurun_out/r02_bench_under_ncu.log 2>&1
