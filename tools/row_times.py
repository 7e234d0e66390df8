"""Per-row time stamps of the merge scan (LIMGCU_MERGE_ROWTIMES=1): prints the wavefront's slope and where rows spend time."""
import os, sys
os.environ["LIMGCU_MERGE_ROWTIMES"] = os.environ.get("ROW", "120")
os.environ.setdefault("LIMGCU_PLAN_ASYNC", "0")
sys.path.insert(0, ".")
import numpy as np, torch
from limg_b200 import Codec, synth
c = Codec(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c2_4k_photo"
img, alpha = synth.CONFIGS[name]()
h, w = img.shape
d = torch.from_numpy(img.view(np.int32)).cuda()
codes = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(3)]
stream = {"codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
for it in range(3):
    c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, stream, None)
    c.sync()
by = (h + 7) // 8
t = c.debug_wave_rows(by).astype(np.int64)
for st in range(2):
    r = t[st]
    t0 = t[0][:, 0].min()
    r = (r - t0) / 1e3
    first, last, done = r[:, 1], r[:, 2], r[:, 3]
    has = t[st][:, 1] != 0
    print("stage %d: rows with decisions %d of %d; kernel span %.0f us" % (st, has.sum(), by, r[:, 3].max()))
    ys = np.nonzero(has)[0]
    for y in ys[:: max(1, len(ys) // 24)]:
        print("   row %4d: ticket %8.1f first %8.1f last %8.1f done %8.1f  (row busy %.1f us)" % (y, r[y, 0], first[y], last[y], done[y], last[y] - first[y]))
    if len(ys) > 2:
        print("   slope of 'last decision' vs row: %.2f us/row ; mean row busy time %.1f us" % (np.polyfit(ys, last[ys], 1)[0], (last[ys] - first[ys]).mean()))

ev = c.wave_events.astype(np.int64)
t0 = t[0][:, 0].min()
print("decisions of four consecutive stage-0 rows (column, time us, c = claimed):")
for r in range(4):
    e = ev[r]
    n = int((e[:, 1] != 0).sum())
    print("  row +%d: " % r + " ".join("%d%s@%.0f" % (e[i, 0] & 0xFFFF, "c" if (e[i, 0] >> 16) & 1 else "", (e[i, 1] - t0) / 1e3) for i in range(n)))
