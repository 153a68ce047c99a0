export SHARE_MODES=separate
for poll in 8 16 32 64; do for lim in 128 512 2048; do
echo "poll $poll limit $lim: $(DODRT_DONATE_POLL=$poll DODRT_HELPER_LIMIT=$lim timeout 100 python tests/tools/share_probe.py dragon4k 8 1 2>&1 | tail -1 | sed 's/.*separate//')"
done; done
