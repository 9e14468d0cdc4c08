"""GPU (-m gpu): the CUDA path, called through the C ABI (libvtts.so via ctypes), against
 (1) fixtures produced by the unmodified reference (tests/golden), (2) the oracle on fresh seeded inputs,
 (3) size-independent properties at BASELINE.json's full sizes.
Tolerances: waveform max-abs <= 1e-3 (north_star, fp32); durations / alignment indices bit-exact."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu
WAV_TOL = 1e-3      # north_star tolerance
WAV_TIGHT = 8e-5    # what the engine achieves on a +-0.9 waveform (fp32 FFMA ~2e-6, split-bf16 tensor path ~2.5e-5)


def _case(g, u):
    p = "u%d_" % u
    return dict(tok=g[p + "tokens"], sid=int(g[p + "sid"]), eps_dp=g[p + "eps_dp"], eps_z=g[p + "eps_z"],
                w_ceil=g[p + "w_ceil"], idx=g[p + "idx"], Ty=int(g[p + "y_length"]), wav=g[p + "wav"], z=g[p + "z"],
                z_p=g[p + "z_p"])


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if c != "ragged3"])
def test_golden_single_utterance(engine, name):
    g = load_golden(name)
    c = _case(g, 0)
    T = len(c["tok"])
    engine.debug_flags(1)
    ylen, dur = engine.durations(c["tok"][None], [T], [c["sid"]], g["scales"], c["eps_dp"][None], want_durations=True)
    assert int(ylen[0]) == c["Ty"]
    assert np.array_equal(dur[0], c["w_ceil"]), "durations (w_ceil) must be bit-exact"
    wav, idx = engine.synthesize(ylen, c["eps_z"][None], want_alignment=True)
    assert np.array_equal(idx[0, : c["Ty"]], c["idx"]), "alignment indices must be bit-exact"
    z_p = engine.debug_read("z_p").reshape(c["Ty"], -1)
    z = engine.debug_read("z").reshape(c["Ty"], -1)
    assert np.abs(z_p - c["z_p"].T).max() < 3e-4      # |z_p| ~ 10-16: 2e-5 relative (split-bf16 encoder in mode 2)
    assert np.abs(z - c["z"].T).max() < 3e-4
    err = np.abs(wav[0, : c["Ty"] * 256] - c["wav"]).max()
    assert err < WAV_TOL
    assert err < WAV_TIGHT, "fp32 path drifted: %g" % err
    engine.debug_flags(0)


def test_golden_ragged_batch_in_one_call(engine):
    """Three utterances of different length in ONE call must each equal the reference's B=1 result
    (per-utterance zero halos in the unmasked decoder, SURVEY.md section 7 'ragged batches')."""
    g = load_golden("ragged3")
    n = int(g["n"])
    cs = [_case(g, u) for u in range(n)]
    Tm = max(len(c["tok"]) for c in cs)
    ids = np.full((n, Tm), 61, np.int64)          # garbage beyond the lengths must be ignored
    eps_dp = np.full((n, 2, Tm), 9.0, np.float32)
    for u, c in enumerate(cs):
        ids[u, : len(c["tok"])] = c["tok"]
        eps_dp[u, :, : len(c["tok"])] = c["eps_dp"]
    lens = [len(c["tok"]) for c in cs]
    ylen, dur = engine.durations(ids, lens, [c["sid"] for c in cs], g["scales"], eps_dp, want_durations=True)
    assert [int(v) for v in ylen] == [c["Ty"] for c in cs]
    eps_z = np.full((n, 192, int(ylen.max())), 7.0, np.float32)
    for u, c in enumerate(cs):
        assert np.array_equal(dur[u, : lens[u]], c["w_ceil"])
        eps_z[u, :, : c["Ty"]] = c["eps_z"]
    wav, idx = engine.synthesize(ylen, eps_z, want_alignment=True)
    for u, c in enumerate(cs):
        assert np.array_equal(idx[u, : c["Ty"]], c["idx"])
        assert np.abs(wav[u, : c["Ty"] * 256] - c["wav"]).max() < WAV_TIGHT
        assert not wav[u, c["Ty"] * 256:].any()


@pytest.mark.parametrize("T,seed,sid,scales", [(64, 11, 0, (0.8, 1.0, 0.8)), (200, 12, 4, (0.667, 0.9, 0.8)),
                                               (256, 13, 150, (1.0, 1.1, 1.0))])
def test_fresh_inputs_vs_oracle(engine, folded, cfg, T, seed, sid, scales):
    from oracle import vits_oracle as vo
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, cfg["n_vocab"], (1, T), generator=g)
    eps_dp = torch.randn(1, 2, T, generator=g)
    eps_z = torch.randn(1, 192, 24 * T, generator=g)
    with torch.no_grad():
        o = vo.infer(folded, cfg, tok, torch.tensor([T]), torch.tensor([sid]), scales, eps_dp, eps_z, return_all=True)
    Ty = int(o["y_lengths"][0])
    margin = float(np.abs((torch.exp(o["logw"]) * scales[1]).numpy() - np.round((torch.exp(o["logw"]) * scales[1]).numpy())).min())
    ylen, dur = engine.durations(tok.numpy(), [T], [sid], scales, eps_dp.numpy(), want_durations=True)
    same = np.array_equal(dur[0], o["w_ceil"][0, 0].numpy().astype(np.int32))
    if not same and margin < 1e-4:
        pytest.skip("a duration sits within %.1e of an integer: ceil() may legitimately flip (SURVEY.md section 7)" % margin)
    assert same
    wav = engine.synthesize(ylen, eps_z[:, :, :Ty].numpy())
    assert np.abs(wav[0, : Ty * 256] - o["o"][0, 0].numpy()).max() < WAV_TIGHT


def _rand_batch(cfg, B, lo, hi, seed):
    rng = np.random.RandomState(seed)
    lens = rng.randint(lo, hi + 1, size=B)
    ids = rng.randint(0, cfg["n_vocab"], size=(B, int(lens.max()))).astype(np.int64)
    sid = rng.randint(0, 5, size=B).astype(np.int64)
    return ids, lens, sid


def test_full_size_batch_invariance_and_determinism(engine, cfg):
    """BASELINE.json configs[2] shape (batch 64, 64-256 phonemes): an utterance's samples do not depend on what
    else is in the batch (to summation-order noise: the split-K decomposition of a conv is chosen from the size of
    the launch), and the Philox path is deterministic in its seed (bit-exact run to run)."""
    ids, lens, sid = _rand_batch(cfg, 64, 64, 256, 1)
    scales = (0.8, 1.0, 0.8)
    wav, ylen = engine.infer(ids, lens, sid, scales, seed=42)
    wav2, ylen2 = engine.infer(ids, lens, sid, scales, seed=42)
    assert np.array_equal(ylen, ylen2) and np.array_equal(wav, wav2)
    assert np.isfinite(wav).all() and 0.05 < np.abs(wav).max() < 20
    for b in (0, 17, 63):
        Ty = int(ylen[b])
        assert 64 <= Ty
        assert not wav[b, Ty * 256:].any()
    wav3, ylen3 = engine.infer(ids, lens, sid, scales, seed=43)
    assert not np.array_equal(wav3[:, :1024], wav[:, :1024])
    # noise-free: (noise scales 0) -> independent of the seed, and batch-invariant
    wa, ya = engine.infer(ids, lens, sid, (0.0, 1.0, 0.0), seed=1)
    for b in (3, 40):
        wb, yb = engine.infer(ids[b:b + 1, : lens[b]], lens[b:b + 1], sid[b:b + 1], (0.0, 1.0, 0.0), seed=99)
        assert int(yb[0]) == int(ya[b])
        assert np.abs(wb[0] - wa[b, : int(yb[0]) * 256]).max() < 2e-5


def test_length_scale_scales_durations(engine, cfg):
    ids, lens, sid = _rand_batch(cfg, 4, 100, 128, 5)
    _, d1 = engine.durations(ids, lens, sid, (0.0, 1.0, 0.0), want_durations=True)
    _, d2 = engine.durations(ids, lens, sid, (0.0, 2.0, 0.0), want_durations=True)
    assert (d2 >= d1).all() and (d2 <= 2 * d1).all() and d2.sum() > 1.5 * d1.sum()


def test_long_utterance_2000_phonemes(engine, folded, cfg):
    """BASELINE.json configs[4] length (monolithic here): finite, right size, and the first second equals the oracle
    run on the same inputs only where the oracle is cheap -- so compare durations only."""
    rng = np.random.RandomState(9)
    ids = rng.randint(0, cfg["n_vocab"], size=(1, 2000)).astype(np.int64)
    wav, ylen = engine.infer(ids, [2000], [2], (0.8, 1.0, 0.8), seed=3)
    assert wav.shape[1] == int(ylen[0]) * 256 and np.isfinite(wav).all()


def test_chunked_vocoder_equals_monolithic(engine, cfg):
    """BASELINE.json configs[4] mechanism: flow once, vocode 64-frame chunks with a 24-frame halo that is discarded;
    the concatenation equals the monolithic waveform (the halo covers the decoder's receptive field)."""
    rng = np.random.RandomState(4)
    ids = rng.randint(0, cfg["n_vocab"], size=(1, 300)).astype(np.int64)
    mono, yl = engine.infer(ids, [300], [2], (0.8, 1.0, 0.8), seed=9)
    chunks = list(engine.synthesize_stream(ids, 2, (0.8, 1.0, 0.8), chunk_frames=64, seed=9))
    wav = np.concatenate(chunks)
    assert len(chunks) == -(-int(yl[0]) // 64) and wav.size == mono.shape[1]
    tol = 2e-6 if engine.precision == 0 else 5e-5
    assert np.abs(wav - mono[0]).max() < tol


def test_error_paths(engine, cfg):
    from vosk_tts_b200.engine import VttsError
    ids, lens, sid = _rand_batch(cfg, 2, 8, 16, 2)
    with pytest.raises(VttsError) as e:
        engine.synthesize(np.array([4, 4]))
    assert e.value.code == -5
    with pytest.raises(VttsError) as e:
        engine.durations(ids, [0, 5], sid, (0.8, 1, 0.8))
    assert e.value.code == -1
    ylen = engine.durations(ids, lens, sid, (0.8, 1, 0.8))
    with pytest.raises(VttsError) as e:
        engine.synthesize(ylen, np.zeros((2, 192, 1), np.float32))      # too few noise columns
    assert e.value.code == -4


def test_graph_replay_equals_eager(engine, cfg):
    """The same call three times: eager, graph capture, graph replay -- bit-identical waveforms."""
    ids, lens, sid = _rand_batch(cfg, 2, 40, 60, 21)
    engine.set_graphs(False)
    ref, yl = engine.infer(ids, lens, sid, (0.8, 1.0, 0.8), seed=5)
    engine.set_graphs(True)
    n0 = engine.graph_replays()
    for _ in range(3):
        w, y2 = engine.infer(ids, lens, sid, (0.8, 1.0, 0.8), seed=5)
        assert np.array_equal(y2, yl) and np.array_equal(w, ref)
    assert engine.graph_replays() > n0


def test_one_handle_called_from_several_threads(engine, cfg):
    """The gRPC server of the reference calls ONE shared session from a thread pool (server/tts_server.py:35,57).  Both the
    one-shot call (vtts_infer) and the two-phase pair (vtts_durations -> vtts_synthesize, which keeps state in the handle
    between the calls) must give every thread its own utterance's result."""
    import threading
    g = torch.Generator().manual_seed(5)
    jobs = []
    for i in range(6):
        T = int(torch.randint(20, 90, (1,), generator=g))
        jobs.append([torch.randint(0, cfg["n_vocab"], (1, T), generator=g).numpy(), T, i % 5,
                     torch.randn(1, 2, T, generator=g).numpy(), None])
    serial = []
    for job in jobs:
        tok, T, sid, e1, _ = job
        yl = engine.durations(tok, [T], [sid], (0.8, 1.0, 0.8), e1)
        job[4] = torch.randn(1, 192, int(yl[0]), generator=g).numpy()       # (the frame count is data dependent)
        serial.append(engine.synthesize(yl, job[4]).copy())
    out, errs = [None] * len(jobs), []

    def work(k, two_phase):
        try:
            for rep in range(4):
                tok, T, sid, e1, e2 = jobs[k]
                if two_phase:
                    yl = engine.lib_durations_threadsafe(tok, [T], [sid], (0.8, 1.0, 0.8), e1)
                    w = engine.lib_synthesize_threadsafe(1, yl, e2[:, :, : int(yl[0])])
                else:
                    w, yl = engine.infer(tok, [T], [sid], (0.8, 1.0, 0.8), e1, e2[:, :, : serial[k].shape[1] // 256], frames_hint=serial[k].shape[1] // 256)
                out[k] = np.array(w)
        except Exception as ex:      # noqa: BLE001
            errs.append(repr(ex))

    th = [threading.Thread(target=work, args=(k, k % 2 == 0)) for k in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in range(len(jobs)):
        # (not bit-identical: a one-shot call may run its second phase for a larger predicted length bucket, whose launches
        #  split their k-loops differently -- same arithmetic, another summation order; other utterances differ by ~0.1)
        assert out[k].shape == serial[k].shape and np.abs(out[k] - serial[k]).max() < 2e-5, "thread %d got another utterance's result" % k


def test_speculative_second_phase_hits_and_misses(engine, cfg):
    """Single-utterance infer calls enqueue phase 2 for a predicted length bucket before the durations are known.  Whatever
    the prediction, the result must equal the two-phase API's; a call whose frame count exceeds the prediction (here: a
    much smaller length_scale after a run of normal ones, so that the one-frame-per-token floor dominates) is repeated."""
    g = torch.Generator().manual_seed(91)
    h0, m0 = engine.speculation_stats()
    plan = [(100, 1.0), (90, 1.0), (110, 1.1), (100, 0.25), (96, 1.0), (100, 2.0), (64, 1.0)]
    for T, ls in plan:
        tok = torch.randint(0, cfg["n_vocab"], (1, T), generator=g).numpy()
        eps_dp = torch.randn(1, 2, T, generator=g).numpy()
        scales = (0.8, ls, 0.8)
        yl = engine.durations(tok, [T], [1], scales, eps_dp)
        Ty = int(yl[0])
        eps_z = torch.randn(1, 192, Ty, generator=g).numpy()
        ref = engine.synthesize(yl, eps_z)
        wav, yl2 = engine.infer(tok, [T], [1], scales, eps_dp, eps_z, frames_hint=Ty + 40)
        assert int(yl2[0]) == Ty
        assert np.abs(wav[:, : Ty * 256] - ref[:, : Ty * 256]).max() < 2e-5      # (another bucket => another summation order)
    h1, m1 = engine.speculation_stats()
    assert h1 - h0 >= 4 and m1 - m0 >= 1, (h1 - h0, m1 - m0)


def test_bucketed_graphs_serve_unseen_utterances(engine, cfg):
    """Graphs are captured per LENGTH BUCKET: utterances that were never seen before (other tokens, other lengths inside the
    bucket, other noise) must replay a captured graph and give exactly what eager launches give -- including a short
    utterance right after a longer one of the same bucket (rows behind the utterance's end hold the previous call's data
    until zero_tails_kernel clears them)."""
    g = torch.Generator().manual_seed(77)
    cases = []
    for T in (128, 121, 113, 126, 115, 128, 119):                       # one token bucket (113..128)
        tok = torch.randint(0, cfg["n_vocab"], (1, T), generator=g).numpy()
        eps_dp = torch.randn(1, 2, T, generator=g).numpy()
        eps_z = torch.randn(1, 192, 12 * T, generator=g).numpy()
        cases.append((tok, T, eps_dp, eps_z))
    engine.set_graphs(False)
    refs = []
    for tok, T, eps_dp, eps_z in cases:
        yl = engine.durations(tok, [T], [3], (0.8, 1.0, 0.8), eps_dp)
        refs.append((int(yl[0]), engine.synthesize(yl, eps_z[:, :, : int(yl[0])]).copy()))
    engine.set_graphs(True)
    n0 = engine.graph_replays()
    for rep in range(3):
        for (tok, T, eps_dp, eps_z), (Ty, ref) in zip(cases, refs):
            yl = engine.durations(tok, [T], [3], (0.8, 1.0, 0.8), eps_dp)
            assert int(yl[0]) == Ty
            wav = engine.synthesize(yl, eps_z[:, :, :Ty])
            assert np.array_equal(wav, ref), "graph replay differs from eager launches (rep %d, T=%d)" % (rep, T)
    # after the buckets have been seen twice everything replays: 7 utterances x 2 phases in the last repetition alone
    assert engine.graph_replays() - n0 >= 14


@pytest.mark.parametrize("env", [{"VTTS_TC_BN": "128"}, {"VTTS_TC_TALL": "1"}, {"VTTS_PDL": "0"}, {"VTTS_CONV_MAXS": "1", "VTTS_CONV_MAXG": "4"},
                                 {"VTTS_ATTN_ROWS": "4"}, {"VTTS_TC_MULTICAST": "1"}, {"VTTS_TC_SPLIT": "1"}, {"VTTS_TC_SPLIT": "2"},
                                 {"VTTS_TC_MINSTEPS": "1"}, {"VTTS_TC_BN": "128", "VTTS_TC_MINSTEPS": "1"}, {"VTTS_TC_BN": "64"}, {"VTTS_ATTN_SPLIT": "0"},
                                 {"VTTS_CONV_AUTOG": "0"}, {"VTTS_MRF_BRANCH": "1"}, {"VTTS_BUCKETS": "0"}, {"VTTS_ATTN_TC": "0"}],
                         ids=["tc-128-wide-tiles", "tc-tall-activation-tiles", "no-programmatic-dependent-launch", "ffma-no-cluster-4-groups",
                              "attention-4-rows-per-warp", "tc-tma-multicast-cluster", "tc-no-split-k", "tc-split-k-pairs",
                              "tc-split-k-8-ways", "tc-128-wide-split-k-8-ways", "tc-64-wide-only", "attention-without-split-kv",
                              "ffma-single-thread-group", "mrf-chains-on-separate-streams", "exact-sizes-no-length-buckets",
                              "ffma-attention-in-the-flow"])
def test_alternative_kernel_configurations_match_golden(packed, cfg, env):
    """The tuning switches select different tilings / launch modes of the same kernels (128-wide tcgen05 tiles are what
    batched calls use automatically); each must still reproduce the reference fixture."""
    import os
    from vosk_tts_b200.engine import Engine
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        e = Engine(cfg, packed[0], packed[1], device=0, precision=1)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for name in ("t128_sid2", "t17_sid2"):
        g = load_golden(name)
        c = _case(g, 0)
        T = len(c["tok"])
        for rep in range(3):          # eager, capture, replay
            ylen, dur = e.durations(c["tok"][None], [T], [c["sid"]], g["scales"], c["eps_dp"][None], want_durations=True)
            assert np.array_equal(dur[0], c["w_ceil"])
            wav = e.synthesize(ylen, c["eps_z"][None])
            assert np.abs(wav[0, : c["Ty"] * 256] - c["wav"]).max() < WAV_TIGHT
    e.close()


def test_plain_hifigan_resblock2_variant_vs_oracle(cfg):
    """Config-driven variant: plain HiFi-GAN `Generator` tail (conv_post + tanh, speaker projection added to conv_pre,
    models.py:845-898) with ResBlock2 (modules.py:234-258), three upsampling stages 8x8x4.  The reference's own `infer`
    cannot drive this decoder (`o, o_mb = self.dec(...)` at models.py:1703 fails to unpack the single tensor it returns),
    so the engine is compared with the oracle restatement only (fp32 path)."""
    import copy
    from oracle import vits_oracle as vo
    from vosk_tts_b200 import synthetic, weights
    from vosk_tts_b200.engine import Engine
    c2 = copy.deepcopy(cfg)
    c2.update(decoder="hifigan", resblock="2", resblock_kernel_sizes=[3, 5, 7], resblock_dilation_sizes=[[1, 2], [2, 6], [3, 9]],
              upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8], upsample_initial_channel=256, n_layers=3)
    folded = weights.fold_weight_norm(synthetic.make_random_checkpoint(c2, 77))
    blob, man = weights.pack(folded, c2)
    e = Engine(c2, blob, man, device=0, precision=0)
    assert e.hop == 256
    g = torch.Generator().manual_seed(5)
    T = 40
    tok = torch.randint(0, 62, (1, T), generator=g)
    e1, e2 = torch.randn(1, 2, T, generator=g), torch.randn(1, 192, 24 * T, generator=g)
    with torch.no_grad():
        o = vo.infer(folded, c2, tok, torch.tensor([T]), torch.tensor([4]), (0.8, 1.0, 0.8), e1, e2)
    ylen, dur = e.durations(tok.numpy(), [T], [4], (0.8, 1.0, 0.8), e1.numpy(), want_durations=True)
    assert np.array_equal(dur[0], o["w_ceil"][0, 0].numpy().astype(np.int32))
    Ty = int(ylen[0])
    wav = e.synthesize(ylen, e2[:, :, :Ty].numpy())
    assert np.abs(wav[0, : Ty * 256] - o["o"][0, 0].numpy()).max() < 2e-5
    e.close()


def test_fused_infer_equals_two_phase_and_handles_capacity(engine, cfg):
    """vtts_infer (one ABI call) == vtts_durations + vtts_synthesize; a too-small capacity falls back cleanly."""
    ids, lens, sid = _rand_batch(cfg, 3, 20, 45, 31)
    ref, yl = engine.infer(ids, lens, sid, (0.8, 1.0, 0.8), seed=11)
    w1, y1 = engine.infer(ids, lens, sid, (0.8, 1.0, 0.8), seed=11, frames_hint=int(yl.max()) + 7)
    assert np.array_equal(y1, yl) and np.array_equal(w1[:, : ref.shape[1]], ref)
    w2, y2 = engine.infer(ids, lens, sid, (0.8, 1.0, 0.8), seed=11, frames_hint=max(1, int(yl.max()) // 2))   # CAPACITY path
    assert np.array_equal(y2, yl) and np.array_equal(w2, ref)


def test_session_run_matches_reference_call_shape(packed, cfg):
    from vosk_tts_b200.session import VitsSession
    s = VitsSession(cfg=cfg, packed=packed, device=0, seed=7, precision=1)
    ids = np.random.RandomState(0).randint(0, 62, size=(1, 40)).astype(np.int64)
    feeds = {"input": ids, "input_lengths": np.array([40], np.int64), "scales": np.array([0.8, 1.0, 0.8], np.float32),
             "sid": np.array([2], np.int64), "bert": None, "phone_duration_extra": None}
    out = s.run(None, feeds)[0]
    assert out.dtype == np.float32 and out.ndim == 4 and out.shape[:3] == (1, 1, 1)
    assert out.shape[3] == int(s.last_y_lengths[0]) * 256
    s.close()


@pytest.mark.parametrize("precision", [0, 1])
def test_plain_coupling_flow_variant_vs_reference_fixture(precision):
    """use_transformer_flows=False (plain ResidualCouplingLayer + Flip == ResidualCouplingBlock, models.py:765-810,
    vc/modules.py:300-345) against the fixture the unmodified reference produced with that flag."""
    import copy
    import json
    from vosk_tts_b200 import config as C, synthetic, weights
    from vosk_tts_b200.engine import Engine
    g = load_golden("plainflow_t40")
    cfg = copy.deepcopy(C.DEFAULT_CONFIG)
    cfg.update(json.loads(str(g["model_overrides"])))
    w = weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, int(g["weight_seed"])))
    blob, man = weights.pack(w, cfg)
    eng = Engine(cfg, blob, man, device=0, precision=precision)
    try:
        c = _case(g, 0)
        T = len(c["tok"])
        eng.debug_flags(1)
        ylen, dur = eng.durations(c["tok"][None], [T], [c["sid"]], g["scales"], c["eps_dp"][None], want_durations=True)
        assert int(ylen[0]) == c["Ty"]
        assert np.array_equal(dur[0], c["w_ceil"])
        wav, idx = eng.synthesize(ylen, c["eps_z"][None], want_alignment=True)
        assert np.array_equal(idx[0, : c["Ty"]], c["idx"])
        z = eng.debug_read("z").reshape(c["Ty"], -1)
        assert np.abs(z - c["z"].T).max() < 3e-4
        assert np.abs(wav[0, : c["Ty"] * 256] - c["wav"]).max() < WAV_TIGHT
    finally:
        eng.close()
