"""CPU tests that pin the SBR part of the oracle (C++ restatement of JAAD's sbr package).

  1. integer stage: header -> band tables, grid, delta-decoded envelopes and noise floors against the generator's own
     integer model (gen/aacgen_sbr.inc: independent band-table derivation and delta arithmetic, no decoder involved);
  2. QMF banks: JAAD's 32-band analysis and 64-band synthesis against float64 direct-form evaluations of the
     ISO 14496-3 4.6.18.4 definitions;
  3. drift: committed fixtures (tests/golden/sbr_*.npz).
"""
import hashlib
import os
import re

import numpy as np
import pytest

import gen
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sbr_cfg(mono, n_frames):
    if mono:
        return gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=n_frames, target_bytes=171, sbr_mode=1)
    return gen.config(3, n_frames=n_frames)


@pytest.mark.parametrize("mono", [False, True], ids=["stereo", "mono"])
def test_sbr_integer_stage_matches_generator_truth(mono):
    n_checked = 0
    classes, couplings, resets = set(), set(), 0
    for seed in range(24):
        cfg = sbr_cfg(mono, 64)   # headers at frames 0, 20, 40, 60 -- some with new contents (decoder reset)
        cfg.sbr_quirk = bool(seed & 1)   # odd seeds: coupled frames in the form only the reference parses
        st = gen.generate(cfg, 31000 + seed, with_truth=True)
        dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
        for f in range(cfg.n_frames):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0, (seed, f, r["status"])
            assert (r["channels"], r["sample_length"], r["sample_rate"]) == (2, 2048, 48000)
            assert np.abs(r["f32"]).max() < 32768.0   # the generator keeps PCM inside int16
            for ch in range(1 if mono else 2):
                t = dec.tap_sbr(0, ch)
                tr = st.truth["sbr"][f, ch]
                L_E, L_Q = int(t["ints"][0]), int(t["ints"][1])
                assert np.array_equal(t["ints"][:3], tr[:3]), (seed, f, ch)
                if t["ints"][2] != 0:
                    assert t["ints"][3] == tr[3]                                    # bs_pointer (not sent for FIXFIX)
                assert np.array_equal(t["ints"][4:5 + L_E], tr[4:5 + L_E])           # t_E
                assert np.array_equal(t["ints"][10:10 + L_E], tr[10:10 + L_E])       # f
                assert np.array_equal(t["ints"][16:18], tr[16:18])                   # amp_res, coupling
                assert np.array_equal(t["extra"][:7], tr[18:25]), (seed, f, ch)      # kx, M, N_high, N_low, N_Q, k0, N_master
                E_o, E_t = t["ints"][32:352].reshape(5, 64), tr[32:352].reshape(5, 64)
                Q_o, Q_t = t["ints"][352:480].reshape(2, 64), tr[352:480].reshape(2, 64)
                for l in range(L_E):
                    nb = t["extra"][2] if t["ints"][10 + l] else t["extra"][3]
                    assert np.array_equal(E_o[l, :nb], E_t[l, :nb]), (seed, f, ch, l)
                assert np.array_equal(Q_o[:L_Q, :t["extra"][4]], Q_t[:L_Q, :t["extra"][4]]), (seed, f, ch)
                classes.add(int(t["ints"][2]))
                couplings.add(int(t["ints"][17]))
                resets += int(t["extra"][8]) if ch == 0 and f > 0 else 0
                n_checked += 1
    assert n_checked == 24 * 64 * (1 if mono else 2)
    assert classes == {0, 1, 2, 3}
    assert couplings == ({0} if mono else {0, 1})
    assert resets >= 1   # at least one mid-stream header change made the decoder rebuild its tables


def test_ps_integer_stage_matches_generator_truth():
    """HE-AAC v2: the parametric-stereo parameters the oracle reconstructs (envelope count and borders, IID / ICC indices
    after delta decoding, clipping and stride expansion) against the generator's own index arithmetic."""
    modes = set()
    n_checked = 0
    for seed in range(24):
        cfg = gen.config(4, n_frames=40)
        st = gen.generate(cfg, 52000 + seed, with_truth=True)
        dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
        stereo_seen = False
        for f in range(cfg.n_frames):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0, (seed, f, r["status"])
            assert (r["channels"], r["sample_length"], r["sample_rate"]) == (2, 2048, 48000)
            stereo_seen |= not np.array_equal(r["f32"][0], r["f32"][1])
            t, tr = dec.tap_ps(0), st.truth["ps"][f]
            ne = t["num_env"]
            assert ne == tr[0] and np.array_equal(t["border"][:ne + 1], tr[1:2 + ne]), (seed, f)
            assert np.array_equal(t["iid"][:ne, :20], tr[8:178].reshape(5, 34)[:ne, :20]), (seed, f)
            assert np.array_equal(t["icc"][:ne, :20], tr[178:348].reshape(5, 34)[:ne, :20]), (seed, f)
            modes.add((t["iid_mode"], t["icc_mode"]))
            n_checked += 1
        assert stereo_seen
    assert n_checked == 24 * 40
    assert len(modes) >= 6   # default / fine IID, type-A / type-B mixing, 10- and 20-band, disabled


def _qmf_c():
    txt = open(os.path.join(ROOT, "jaadec_b200", "csrc", "generated", "jaad_tables.h")).read()
    m = re.search(r'JAAD_TABLE_F32\(SBR_QMF_C, 640, "\[640\]"\)(.*?)JAAD_TABLE_END', txt, re.S)
    bits = np.array([int(t, 16) for t in re.findall(r"0x([0-9A-Fa-f]{8})u", m.group(1))], np.uint32)
    assert len(bits) == 640
    return bits.view(np.float32).astype(np.float64)


def test_qmf_banks_match_direct_form():
    rng = np.random.default_rng(7)
    n = 3
    x = (rng.standard_normal(n * 1024) * 2000).astype(np.float32)
    X, pcm = oracle.qmf_roundtrip(x)
    c = _qmf_c()
    xx = np.concatenate([np.zeros(320), x.astype(np.float64)])
    nn, kk = np.arange(64), np.arange(32)
    A = 2 * np.exp(1j * np.pi / 64 * (kk[:, None] + 0.5) * (2 * nn[None, :] - 0.5))      # 4.6.18.4.1
    Xd = np.zeros((n * 32, 32), complex)
    for l in range(n * 32):
        newest = 320 + 32 * l + 31
        u = (xx[newest - np.arange(320)] * c[::2]).reshape(5, 64).sum(0)
        Xd[l] = A @ u
    assert np.abs(X - Xd).max() / np.abs(Xd).max() < 2e-6
    n128, k64 = np.arange(128), np.arange(64)
    S = np.exp(1j * np.pi / 128 * (k64[None, :] + 0.5) * (2 * n128[:, None] - 255)) / 64.0   # 4.6.18.4.2
    v = np.zeros(1280)
    out = np.zeros(n * 2048)
    for l in range(n * 32):
        Xl = np.zeros(64, complex)
        Xl[:32] = Xd[l]
        v = np.concatenate([(S @ Xl).real, v[:-128]])
        g = np.zeros(640)
        for j in range(5):
            g[128 * j:128 * j + 64] = v[256 * j:256 * j + 64]
            g[128 * j + 64:128 * j + 128] = v[256 * j + 192:256 * j + 256]
        out[64 * l:64 * l + 64] = (g * c).reshape(10, 64).sum(0)
    assert np.abs(out - pcm).max() / np.abs(pcm).max() < 2e-6
    # and the chain reconstructs a low-pass input (x2 up-sampled, delayed): energy is preserved to a fraction of a dB
    lp = np.convolve(rng.standard_normal(8 * 1024), np.ones(8) / 8, "same").astype(np.float32) * 3000
    _, y = oracle.qmf_roundtrip(lp)
    e_in, e_out = float((lp[2048:6144].astype(np.float64) ** 2).mean()), float((y[4096:12288].astype(np.float64) ** 2).mean())
    assert abs(10 * np.log10(e_out / e_in)) < 0.5


def test_downsampled_synthesis_bank_matches_direct_form():
    """sbr/SynthesisFilterbank32.java (DCT4_32 / DST4_32 operation lists from tools/extract_dct32.py) against the
    down-sampled synthesis of ISO 14496-3 4.6.18.4.3 written from its definition in float64."""
    rng = np.random.default_rng(11)
    n = 3
    x = (rng.standard_normal(n * 1024) * 2000).astype(np.float32)
    X, _ = oracle.qmf_roundtrip(x)             # the analysis bank's output (checked above)
    pcm = oracle.qmf_roundtrip32(x)
    c = _qmf_c()
    n64, k32 = np.arange(64), np.arange(32)
    S = np.exp(1j * np.pi / 64 * (k32[None, :] + 0.5) * (2 * n64[:, None] - 127.5)) / 64.0   # the quarter-sample term is qmf32_pre_twiddle
    v = np.zeros(640)
    out = np.zeros(n * 1024)
    for l in range(n * 32):
        v = np.concatenate([(S @ X[l].astype(complex)).real, v[:-64]])
        g = np.zeros(320)
        for j in range(5):
            g[64 * j:64 * j + 32] = v[128 * j:128 * j + 32]
            g[64 * j + 32:64 * j + 64] = v[128 * j + 96:128 * j + 128]
        out[32 * l:32 * l + 32] = (g * c[::2]).reshape(10, 32).sum(0)
    assert np.abs(out - pcm).max() / np.abs(pcm).max() < 2e-6
    # analysis + down-sampled synthesis is a (near-perfect-reconstruction) delay of 289 samples
    assert np.abs(pcm[289:] - x[:-289]).max() / np.abs(x).max() < 2e-3


@pytest.mark.parametrize("name", ["sbr_c3_stereo", "sbr_mono", "ps_c4_mono", "sbr_ds_stereo", "ps_ds_mono", "ps_ipdopd_mono"])
def test_oracle_reproduces_sbr_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n_streams = int(g["frame_stream"].max()) + 1
    asc = g["asc"].tobytes()
    decs = [oracle.Decoder.create_asc(asc) if asc else oracle.Decoder.create_adts(*[int(x) for x in g["hdr"]]) for _ in range(n_streams)]
    sha = hashlib.sha256()
    for i, (o, n, s) in enumerate(zip(g["frame_offset"], g["frame_nbytes"], g["frame_stream"])):
        r = decs[s].decode_frame(g["blob"][o:o + n])
        assert r["status"] == 0
        assert np.array_equal(r["s16"], g["s16"][i]), (name, i)
        sha.update(np.ascontiguousarray(r["f32"], np.float32).tobytes())
    assert sha.digest() == g["f32_sha256"].tobytes()


def test_ps_ipdopd_extension_parse_matches_generator_truth():
    """IPD/OPD extension of parametric stereo (ps/Extension.java, ExtData.java, PDData.java, PDMode.java): the oracle's
    ps_extension parse and modulo-8 delta decoding against the generator's own bookkeeping (no decoder involved), including
    Extension.nr_par -- which decides the parameter bands that get the phase rotation -- and frames whose extension carries
    no phase data (enable_ipdopd = 0: the rotation still runs on the indices an earlier frame left, as in JAAD)."""
    import gen
    import oracle
    n_rot = n_stale = 0
    for seed in range(8):
        cfg = gen.config(4, n_frames=50, ps_ext=0.7)
        st = gen.generate(cfg, gen.seed_for(4, 300 + seed), with_truth=True)
        dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
        for f in range(cfg.n_frames):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0, (seed, f)
            t, tr = dec.tap_ps(0), st.truth["ps"][f]
            ne = t["num_env"]
            assert ne == tr[0] and np.array_equal(t["border"][:ne + 1], tr[1:2 + ne]), (seed, f)
            assert np.array_equal(t["iid"][:ne], tr[8:178].reshape(5, 34)[:ne]) and np.array_equal(t["icc"][:ne], tr[178:348].reshape(5, 34)[:ne])
            assert t["nr_ipdopd_par"] == tr[433], (seed, f)
            assert np.array_equal(t["ipd"], tr[348:433].reshape(5, 17)), (seed, f)
            n_rot += int(tr[433] > 0)
            n_stale += int(tr[433] > 0 and tr[434] == 0)
            assert np.isfinite(r["f32"]).all()
    assert n_rot > 100 and n_stale > 5, (n_rot, n_stale)
