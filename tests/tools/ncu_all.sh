# ncu --set full of the last frame's primary and shadow kernels of every benchmark workload (one gpurun call).
# Each command runs plainly first (&&) as B200_PROFILING.md asks.
set -x
P="python tests/tools/profile_workload.py"
for wl in dragon4k teapot1080 analytic1080 dragon16_8k; do
  $P $wl 3 > gpurun_out/prof_${wl}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 2 -f -o gpurun_out/prof_r02_$wl $P $wl 3 > gpurun_out/prof_${wl}_ncu.log 2>&1
done
wl=dragon1080_primary
$P $wl 3 > gpurun_out/prof_${wl}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_$wl $P $wl 3 > gpurun_out/prof_${wl}_ncu.log 2>&1
