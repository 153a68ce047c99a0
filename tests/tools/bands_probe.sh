# primary-only frame through the host-buffer call (dodrt_trace_frame, pinned buffers): bands of the staged path
for b in 1 2 3 4 6; do
echo "DODRT_BANDS=$b: $(DODRT_BANDS=$b timeout 120 python tests/tools/e2e_probe.py dragon1080_primary 1 2>&1 | grep 'N=1' | sed 's/.*staged \([0-9.]*\) ms.*/\1 ms/')"
done
