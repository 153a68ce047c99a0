# frame shares traced in 1..4 chunks on concurrent streams: ranks 0,1 of 8 / 4 / 2 of dragon4k and the whole frame
export SHARE_MODES=separate
for ch in 1 2 3 4; do
  echo "== DODRT_FRAME_CHUNKS=$ch"
  DODRT_FRAME_CHUNKS=$ch timeout 300 python tests/tools/share_probe.py dragon4k 8,4,2,1 2 2>&1 | grep -o "N=.*separate.*" | sed 's/block-fused nan (min nan) ms   tile queues nan (min nan) ms  //'
done
