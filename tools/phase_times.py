import sys, time, json
sys.path.insert(0,'.')
import numpy as np, torch
from limg_b200 import Codec, synth
c = Codec(0)
c.enable_phase_timing(True)
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ("c1_512_gradient","c5_1080p_frame0","c2_4k_photo","c4_4k_flatui","c3_8k_rgba")):
    img, alpha = synth.CONFIGS[name]()
    h,w = img.shape
    d = torch.from_numpy(img.view(np.int32)).cuda()
    codes = [torch.empty((h,w),dtype=torch.uint8,device='cuda') for _ in range(3)]
    dec = torch.empty((h,w),dtype=torch.int32,device='cuda')
    stream = {"codesA":codes[0].data_ptr(),"codesB":codes[1].data_ptr(),"codesC":codes[2].data_ptr()}
    for it in range(3):
        t=time.time()
        c.blocked_encode3d_device(d.data_ptr(), w, h, alpha, 100, True, False, stream, {"pDecoded": dec.data_ptr()})
        c.sync(); dt=time.time()-t
    ph = c.phase_ms()
    tot = sum(ph.values())
    cnt = c.debug_counters()
    print("   merged %d areas %d small %d large %d | stage0 exp %d reexp %d polls %d ondemand %d | stage1 exp %d reexp %d polls %d ondemand %d | failed tries / flags %s which 0x%x ext slots %d sym slots %d" % (cnt[0],cnt[1],cnt[2],cnt[3],cnt[8],cnt[9],cnt[10],cnt[11],cnt[12],cnt[13],cnt[14],cnt[15],cnt[24:29],cnt[31],cnt[29],cnt[30]))
    print("   profile kcycles: next %d wait %d expand %d (four-way %d, on-demand %d) claim %d prefetch %d" % tuple(cnt[16:23]))
    dbg = c.debug_wave()
    print("   expansion duration histogram (log2 of cycles/256) stage0 %s stage1 %s" % (dbg[0:16].tolist(), dbg[16:32].tolist()))
    print("   on-demand strips [seed, 4way+bitmap, 4way no bitmap]: n %s kcyc %s | - %s %s | four-way attempts %d, centre mispredicted %d, no bitmap %d, built on the fly %d" % (dbg[32:35].tolist(), dbg[36:39].tolist(), dbg[40:43].tolist(), dbg[44:47].tolist(), dbg[48], dbg[49], dbg[50], dbg[53]))
    print("   strips outside the known part [left, up, right, down]: seed %s 4way+bitmap %s 4way-no-bitmap %s ; known reach (blocks from centre) when right/down left it: %s" % (dbg[64:68].tolist(), dbg[68:72].tolist(), dbg[72:76].tolist(), dbg[80:96].tolist()))
    print("   stage 0 at commit: rows above ahead by (x4 columns) %s ; probe box width (x2) %s ; immediate %d waited %d" % (dbg[128:160].tolist(), dbg[160:192].tolist(), dbg[192], dbg[193]))
    print("   expansion parts (kcycles): seed growth in the 8x8 word %d (%d seeds), general seed growth %d (%d seeds), centre bitmap lookup/build %d, centre region %d, four-way growth %d" % (dbg[224], dbg[230], dbg[225], dbg[231], dbg[226], dbg[227], dbg[228]))
    for st in range(2):
        o = dbg[200 + st * 8: 208 + st * 8].astype(float)
        if o[0] > 0:
            print("   stage %d emitting seeds: %d, per seed (cycles): total %.0f = prefetch %.0f + poll/snapshot %.0f (%.1f iterations) + expand %.0f + claim %.0f + rest" % (st, o[0], 1024 * o[1] / o[0], 1024 * o[2] / o[0], 1024 * o[3] / o[0], o[6] / o[0], 1024 * o[4] / o[0], 1024 * o[5] / o[0]))
    print(name, "%dx%d"%(w,h), "total %.3f ms (wall %.3f) -> %.1f Mpx/s"%(tot, dt*1e3, w*h/tot/1e3), {k: round(v,3) for k,v in ph.items()})
