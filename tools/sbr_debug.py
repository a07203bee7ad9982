#!/usr/bin/env python3
"""Debug aid: one SBR stream, staged decode, compare the engine's SBR frame records with the oracle's state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gen, oracle
from helpers import Workload
from jaadec_b200 import Engine, PCM_F32_PLANAR

seed = int(sys.argv[1]); mono = len(sys.argv) > 2 and sys.argv[2] == "mono"; nfr = int(sys.argv[3]) if len(sys.argv) > 3 else 24
cfg = gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=nfr, target_bytes=171, sbr_mode=1) if mono else gen.config(3, n_frames=nfr)
wl = Workload(cfg, 1, base_seed=seed, with_truth=False)
dec = wl.oracle_decoders()[0]
eng = Engine(max_streams=2, pcm_format=PCM_F32_PLANAR)
ids = [eng.open_adts(*wl.hdr, expect_sbr=1)]
frames, index = wl.frame_table(ids)
b = eng.batch(frames, wl.blob.nbytes); b.upload(wl.blob); b.decode()
pcm, res = b.download()
per = 2 * 2048 * 4
for i, (s, f) in enumerate(index):
    r = dec.decode_frame(wl.frame_bytes(s, f))
    got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
    okp = np.array_equal(got.view(np.uint32), r["f32"].view(np.uint32))
    msgs = []
    for ch in range(1 if mono else 2):
        t = dec.tap_sbr(0, ch); g = b.tap_sbr(i, ch)
        L_E, L_Q = int(t["ints"][0]), int(t["ints"][1])
        ex = t["extra"]
        if (g["L_E"], g["L_Q"], g["kx"], g["M"], g["N_high"], g["N_low"], g["N_Q"], g["noPatches"], g["reset"]) != (L_E, L_Q, ex[0], ex[1], ex[2], ex[3], ex[4], ex[7], ex[8]):
            msgs.append(("hdr", ch, [int(g[k]) for k in ("L_E","L_Q","kx","M","N_high","N_low","N_Q","noPatches","reset")], [L_E, L_Q] + ex[:5].tolist() + [ex[7], ex[8]]))
        if not np.array_equal(g["t_E"][:L_E+1], t["ints"][4:5+L_E]): msgs.append(("t_E", ch, g["t_E"], t["ints"][4:10]))
        for l in range(L_E):
            nb = ex[2] if t["ints"][10+l] else ex[3]
            if not np.array_equal(g["E_orig"][l,:nb].view(np.uint32), t["e_orig"][l,:nb].view(np.uint32)): msgs.append(("E_orig", ch, l, g["E_orig"][l,:nb], t["e_orig"][l,:nb]))
        for l in range(L_Q):
            if not np.array_equal(g["Q_div"][l,:ex[4]].view(np.uint32), t["q_div"][l,:ex[4]].view(np.uint32)): msgs.append(("Q_div", ch, l, g["Q_div"][l,:ex[4]], t["q_div"][l,:ex[4]]))
    if not okp or msgs:
        bad = np.argwhere(got.view(np.uint32) != r["f32"].view(np.uint32))
        print("frame", f, "pcm", "ok" if okp else ("BAD n=%d first=%s maxerr=%g" % (len(bad), bad[0].tolist(), np.abs(got - r["f32"]).max())), msgs)
        g = b.tap_sbr(i, 0)
        print("  engine rec:", {k: (g[k].tolist() if hasattr(g[k], "tolist") else g[k]) for k in ("mode","reset","L_E","L_Q","kx","M","N_high","N_low","N_Q","N_L","kx_prev","M_prev","noPatches","limiter_gains","interpol_freq","smoothing_mode","l_A","prevEnvIsShort","t_E","t_Q","f","bs_invf_mode","patchNoSubbands","patchStartSubband","f_table_noise")})
        print("  f_table_lim", g["f_table_lim"][:12].tolist(), "res lo", g["f_table_res"][0][:10].tolist(), "hi", g["f_table_res"][1][:16].tolist())
        print("  oracle extra", t["extra"].tolist(), "ints", t["ints"][:18].tolist())
        break
else:
    print("all frames ok")
