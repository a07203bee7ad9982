#!/usr/bin/env python3
"""Debug aid for tools/fuzz_gpu.py: same mutations, one stream in detail (taps of the first differing frame).

    python tools/fuzz_debug.py <config> <streams> <frames> <seed> <p_corrupt> <stream>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gen, oracle
from helpers import Workload
from jaadec_b200 import Engine, PCM_F32_PLANAR, FLAG_DEBUG_TAPS
cfg_no, n, nf, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
p_corrupt, S = float(sys.argv[5]), int(sys.argv[6])
cfg = gen.config(cfg_no, n_frames=nf, adts=True) if cfg_no == 5 else gen.config(cfg_no, n_frames=nf)
wl = Workload(cfg, n, base_seed=seed, with_truth=False)
rng = np.random.default_rng(seed)
blob = wl.blob.copy()
frames, index = wl.frame_table(list(range(n)))
frames = frames.copy()
mut = {}
for i, (s, f) in enumerate(index):
    if rng.random() < p_corrupt:
        o, nb = int(frames["offset"][i]), int(frames["nbytes"][i])
        kind = rng.integers(0, 4)
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                b = int(rng.integers(0, nb * 8))
                blob[o + b // 8] ^= 1 << (7 - b % 8)
                mut.setdefault((s, f), []).append(("flip", b))
        elif kind == 1:
            frames["nbytes"][i] = int(rng.integers(1, nb)); mut[(s, f)] = [("trunc", int(frames["nbytes"][i]))]
        elif kind == 2:
            a = int(rng.integers(0, nb)); e = min(nb, a + int(rng.integers(1, 16)))
            blob[o + a:o + e] = rng.integers(0, 256, e - a, dtype=np.uint8); mut[(s, f)] = [("burst", a, e)]
        else:
            b = int(rng.integers(0, min(nb, 8) * 8))
            blob[o + b // 8] ^= 1 << (7 - b % 8); mut[(s, f)] = [("hdrflip", b)]
decs = wl.oracle_decoders()
eng = Engine(max_streams=n, pcm_format=PCM_F32_PLANAR, flags=FLAG_DEBUG_TAPS, sbr_tile_frames=int(os.environ.get("TILE", "0")))
ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(n)]
b = eng.batch(frames, blob.nbytes); b.upload(blob); b.decode(); pcm, res = b.download()
info = eng.stream_info(ids[0])
nch, ln = info.channels, info.sample_length
per = nch * ln * 4
shown = 0
if os.environ.get('DUMP'):   # engine output and the mutated input of stream S, for offline comparison against oracle variants
    sel = [i for i, (s, f) in enumerate(index) if s == S]
    np.savez_compressed(os.environ['DUMP'], pcm=np.stack([pcm[i * per:(i + 1) * per].view(np.float32).reshape(nch, ln) for i in sel]),
                        status=res["status"][sel], frames=[blob[int(frames["offset"][i]):int(frames["offset"][i]) + int(frames["nbytes"][i])].tobytes() for i in sel])
for i, (s, f) in enumerate(index):
    if s != S: continue
    o, nb = int(frames["offset"][i]), int(frames["nbytes"][i])
    r = decs[s].decode_frame(blob[o:o + nb])
    line = [f, int(res["status"][i]), r["status"], mut.get((s, f))]
    if r["status"] == 0 and res["status"][i] == 0:
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(nch, ln)
        ref = np.ascontiguousarray(r["f32"], np.float32)
        d = got.view(np.uint32) != ref.view(np.uint32)
        line += [int(d.sum())]
        if d.any() and shown < 2:
            shown += 1
            w = np.argwhere(d)
            line += ["ch with diffs", sorted(set(w[:, 0].tolist())), "first", w[:3].tolist(), [float(got[tuple(x)]) for x in w[:3]], [float(ref[tuple(x)]) for x in w[:3]],
                     "nonfinite got/ref", int((~np.isfinite(got)).sum()), int((~np.isfinite(ref)).sum())]
            el = 0; c = 0
            while True:
                t = decs[s].tap_ics(el, 0)
                if t is None: break
                for k in range(2):
                    t = decs[s].tap_ics(el, k)
                    if t is None: continue
                    g = b.tap(i, c)
                    line += ["slot", c, "q", bool(np.array_equal(g["q"], t["q"])), "sf", bool(np.array_equal(g["sfidx"], t["sfidx"])), "cb", bool(np.array_equal(g["sfbcb"], t["sfbcb"])),
                             "spec", bool(np.array_equal(g["spec"].view(np.uint32), t["spec"].view(np.uint32))), "info", g["info"].tolist(), t["info"].tolist()]
                    c += 1
                el += 1
    if os.environ.get('ALLTAPS'):
        for c in range(cfg.chan_cfg if cfg.chan_cfg <= 2 else 2):
            g = b.tap(i, c, want_spec=False)
            t = decs[s].tap_ics(0, c)
            line += ['tap', c, g['info'][:5].tolist(), t['info'][:5].tolist() if t else None]
    if os.environ.get('SBRTAPS') and r["status"] == 0 and res["status"][i] == 0:
        # engine's SBR frame record vs the oracle's SBR state after this frame
        for ch in range(cfg.chan_cfg):
            t = decs[s].tap_sbr(0, ch); g = b.tap_sbr(i, ch)
            if t is None or g is None: continue
            L_E, L_Q = int(t["ints"][0]), int(t["ints"][1]); ex = t["extra"]
            hdr_g = [int(g[k]) for k in ("L_E", "L_Q", "kx", "M", "N_high", "N_low", "N_Q", "noPatches", "reset")]
            hdr_t = [L_E, L_Q] + ex[:5].tolist() + [int(ex[7]), int(ex[8])]
            if hdr_g != hdr_t: line += ["sbr hdr", ch, hdr_g, hdr_t]
            if not np.array_equal(g["t_E"][:L_E + 1], t["ints"][4:5 + L_E]): line += ["t_E", ch, g["t_E"].tolist(), t["ints"][4:10].tolist()]
            for l in range(min(L_E, 5)):
                nb = ex[2] if t["ints"][10 + l] else ex[3]
                if not np.array_equal(g["E_orig"][l, :nb].view(np.uint32), t["e_orig"][l, :nb].view(np.uint32)): line += ["E_orig differs", ch, l]
            for l in range(min(L_Q, 2)):
                if not np.array_equal(g["Q_div"][l, :ex[4]].view(np.uint32), t["q_div"][l, :ex[4]].view(np.uint32)): line += ["Q_div differs", ch, l]
            line += ["rec", ch, {k: (g[k].tolist() if hasattr(g[k], "tolist") else g[k]) for k in ("mode", "reset", "L_E", "L_Q", "kx", "M", "kx_prev", "M_prev", "l_A", "prevEnvIsShort", "smoothing_mode", "interpol_freq", "limiter_gains", "add_harmonic_flag_prev", "t_E", "f", "bs_invf_mode")},
                     "harm", int(g["bs_add_harmonic"].sum()), int(g["bs_add_harmonic_prev"].sum()), "oracle extra", ex.tolist()]
            if os.environ.get('SBRTAPS') == '2':
                line += ["Q_div", g["Q_div"].tolist(), "Q_div2", g["Q_div2"].tolist(), "lim", g["f_table_lim"][:12].tolist(), "noise", g["f_table_noise"].tolist(),
                         "map", g["table_map_k_to_g"][:40].tolist(), "patches", g["patchNoSubbands"].tolist(), g["patchStartSubband"].tolist(),
                         "res", g["f_table_res"][0][:12].tolist(), g["f_table_res"][1][:14].tolist(), "N_L", int(g["N_L"]), "t_Q", g["t_Q"].tolist(),
                         "E_orig0", g["E_orig"][0][:10].tolist()]
    if os.environ.get('PSTAPS') and r["status"] == 0 and res["status"][i] == 0 and cfg.sbr_mode > 1:
        g, t = b.tap_ps(i), decs[s].tap_ps(0)
        if g is not None and t is not None:
            ne = int(g["num_env"])
            line += ["ps", "use", int(g["use_ps"]), "num_env", ne, int(t["num_env"]), "modes", (int(g["iid_mode"]), int(g["icc_mode"])), (t["iid_mode"], t["icc_mode"]),
                     "border", g["border"][:ne + 1].tolist(), t["border"][:ne + 1].tolist(),
                     "iid eq", bool(np.array_equal(g["iid"][:ne], t["iid"][:ne, :20])), "icc eq", bool(np.array_equal(g["icc"][:ne], t["icc"][:ne, :20]))]
            if not np.array_equal(g["icc"][:ne], t["icc"][:ne, :20]): line += ["icc", g["icc"][:ne].tolist(), t["icc"][:ne, :20].tolist()]
            if not np.array_equal(g["iid"][:ne], t["iid"][:ne, :20]): line += ["iid", g["iid"][:ne].tolist(), t["iid"][:ne, :20].tolist()]
    print(line)
