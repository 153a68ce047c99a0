# sweep of the donation knobs on ranks 0,1 of an 8-way split (and rank 0 of 4) of dragon4k, separate passes
export SHARE_MODES=separate
for cfg in "0 4294967295" "0 512" "8 4294967295" "8 512" "32 4294967295" "32 512" "16 1024"; do
  set -- $cfg
  echo "== fork_poll=$1 helper_limit=$2"
  DODRT_FORK_POLL=$1 DODRT_HELPER_LIMIT=$2 timeout 300 python tests/tools/share_probe.py dragon4k 8,4 2 2>&1 | grep -o "N=.*separate.*" | sed 's/block-fused nan (min nan) ms   tile queues nan (min nan) ms  //'
done
