# in-kernel timeline (globaltimer marks) of the two passes of rank 0 of an 8-way split, debug build -DDODRT_TIMELINE
export DODRT_LIB=$PWD/dod_raytracer_b200/lib/libdodrt_cuda_timeline.so SHARE_MODES=separate
for cfg in "0 512" "8 512" "32 512" "0 4294967295"; do
  set -- $cfg
  echo "== fork_poll=$1 helper_limit=$2"
  DODRT_FORK_POLL=$1 DODRT_HELPER_LIMIT=$2 timeout 300 python tests/tools/share_probe.py dragon4k 8 1 2>&1 | tail -9
done
