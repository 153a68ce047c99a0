# in-kernel timeline of the two passes of rank 0 of an 8-way split (debug build -DDODRT_TIMELINE): globaltimer marks, the
# cycles a resumed ray spends in node / leaf steps, histograms of main-loop exits and donations over 16-us bins.
#   make -C dod_raytracer_b200/csrc timeline   (-> dod_raytracer_b200/lib/libdodrt_cuda_timeline.so; delete it afterwards)
#   TIMELINE_HIST=1 also prints the histograms
export DODRT_LIB=$PWD/dod_raytracer_b200/lib/libdodrt_cuda_timeline.so SHARE_MODES=separate
filter() { if [ -n "$TIMELINE_HIST" ]; then tail -${1}; else grep -E "resumed rays|^timeline|rank" | tail -${2}; fi; }
timeout 300 python tests/tools/share_probe.py dragon4k 8 1 2>&1 | filter 40 5
timeout 300 python tests/tools/share_probe.py dragon1080_primary 1 1 2>&1 | filter 20 3
