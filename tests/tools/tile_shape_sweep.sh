# per-rank share time of an 8-way split for several tile shapes (all 8 ranks one after the other on ONE GPU): with 32x32
# tiles a 3840-wide frame has 120 tiles per row, a multiple of 8, so every rank owns whole tile COLUMNS
export SHARE_MODES=separate
for t in 32x32 56x32 104x32 64x16 56x16; do
echo "tile $t: $(SHARE_TILE=$t timeout 200 python tests/tools/share_probe.py dragon4k 8 8 2>&1 | grep 'N=8' | sed 's/.*separate \([0-9.]*\).*/\1/' | tr '\n' ' ')"
done
