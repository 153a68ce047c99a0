# early donation to dedicated helper warps (DODRT_DEDICATED=every,polls,live; 0 = off): 1-of-8 share and whole frame
export SHARE_MODES=separate
# (the knob DODRT_DEDICATED was removed with the experiment; see profiles/r02_donation_fork.txt section 7)
for cfg in 0 16,3,8 16,2,8 8,3,8 32,3,8 16,3,4 16,4,16 16,1,32 64,3,8; do
echo "dedicated $cfg: 1-of-8 $(DODRT_DEDICATED=$cfg timeout 100 python tests/tools/share_probe.py dragon4k 8 1 2>&1 | tail -1 | sed 's/.*separate//')   frame $(DODRT_DEDICATED=$cfg timeout 100 python tests/tools/share_probe.py dragon4k 1 1 2>&1 | tail -1 | sed 's/.*separate//')"
done
