"""the other inverse-STFT decoders of the reference (Multistream_iSTFT_Generator,
iSTFT_Generator).  The oracle restatement and the weight packing are pinned against the UNMODIFIED reference on CPU
(build container only); the CUDA engine's parity run for them is tests/test_gpu_parity.py::test_istft_decoder_variants_vs_oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from oracle import vits_oracle as vo
from vosk_tts_b200 import config as C, synthetic, weights

pytestmark = pytest.mark.skipif(not rh.available(), reason="needs the reference tree")

N_VOCAB = 40


def _training_json(flag):
    j = copy.deepcopy(rh.load_ref_config())
    m = j["model"]
    m.update(inter_channels=64, hidden_channels=64, filter_channels=128, n_heads=2, n_layers=3, kernel_size=3,
             resblock_kernel_sizes=[3, 5], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]], upsample_rates=[4, 4],
             upsample_initial_channel=64, upsample_kernel_sizes=[16, 16], gin_channels=32,
             mb_istft_vits=False, ms_istft_vits=False, istft_vits=False)
    m[flag] = True
    j["data"]["n_speakers"] = 4
    return j


@pytest.mark.parametrize("flag,kind", [("ms_istft_vits", "ms_istft"), ("istft_vits", "istft"), ("mb_istft_vits", "mb_istft")])
def test_oracle_matches_reference_for_decoder_variant(flag, kind):
    tj = _training_json(flag)
    cfg = C.from_training_json(tj, n_vocab=N_VOCAB)
    assert cfg["decoder"] == kind and C.hop_total(cfg) == (64 if kind == "istft" else 256)
    sd = synthetic.make_random_checkpoint(cfg, 11)
    net = rh.build_reference_model(sd, cfg=tj, n_vocab=N_VOCAB)
    folded = weights.fold_weight_norm(sd)
    g = torch.Generator().manual_seed(3)
    T = 19
    tok = torch.randint(0, N_VOCAB, (1, T), generator=g)
    eps_dp = torch.randn(1, 2, T, generator=g)
    eps_z = torch.randn(1, cfg["inter_channels"], 400 * T, generator=g)     # the random SDP of this seed is slow-spoken
    scales = [0.8, 1.0, 0.8]
    torch.set_num_threads(1)
    r = rh.reference_infer(net, tok, torch.tensor([T]), torch.tensor([2]), scales, eps_dp, lambda s: eps_z[:, :, :s[2]])
    with torch.no_grad():
        o = vo.infer(folded, cfg, tok, torch.tensor([T]), torch.tensor([2]), scales, eps_dp, eps_z, return_all=True)
    attn = r["attn"][0, 0]
    assert np.array_equal(o["w_ceil"][0, 0].numpy().astype(np.int32), attn.sum(0).numpy().astype(np.int32))
    assert np.array_equal(o["idx"][0].numpy(), attn.argmax(1).numpy())
    assert o["o"].shape == r["o"].shape and r["o"].shape[-1] == int(o["y_lengths"][0]) * C.hop_total(cfg)
    assert float((o["o"] - r["o"]).abs().max()) < 1e-5


@pytest.mark.parametrize("flag", ["ms_istft_vits", "istft_vits"])
def test_packed_tail_filter_is_what_the_decoder_applies(flag):
    """pack() hands the CUDA tail kernel one 63-tap filter per band: the learned multistream filter, or a unit impulse."""
    tj = _training_json(flag)
    cfg = C.from_training_json(tj, n_vocab=N_VOCAB)
    folded = weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 11))
    blob, man = weights.pack(folded, cfg, tc=False)
    ent = {}
    for line in man.strip().splitlines():
        parts = line.split()
        ent[parts[0]] = [int(x) for x in parts[1:]]
    off, n = ent["dec.pqmf"][0], ent["dec.pqmf"][1]
    bank = blob[off:off + n].reshape(-1, 63)
    if flag == "ms_istft_vits":
        assert np.array_equal(bank, folded["dec.multistream_conv_post.weight"][0].numpy())
        assert "dec.post.b" in ent            # models.py:1095: this conv_post has a bias
    else:
        assert bank.shape == (1, 63) and bank[0, 31] == 1.0 and np.count_nonzero(bank) == 1


def test_engine_config_accepts_every_istft_decoder():
    """All three inverse-STFT decoders map onto the same engine decoder type (the tail kernel takes the filter bank from
    the blob); the GPU parity runs are tests/test_gpu_parity.py::test_istft_decoder_variants_vs_oracle."""
    from vosk_tts_b200 import engine
    for flag in ("ms_istft_vits", "istft_vits", "mb_istft_vits"):
        cfg = C.from_training_json(_training_json(flag), n_vocab=N_VOCAB)
        cc = engine.make_c_config(cfg)
        assert cc.decoder_type == 0 and cc.subbands == (1 if flag == "istft_vits" else 4)


@pytest.mark.parametrize("flag,kind", [("ms_istft_vits", "ms_istft"), ("istft_vits", "istft")])
def test_variant_is_recognised_in_an_exported_graph(flag, kind, tmp_path):
    """model.onnx of the other decoders (exported with the reference recipe): configuration and every tensor pack() needs."""
    from vosk_tts_b200 import onnx_weights as ow
    tj = _training_json(flag)
    cfg = C.from_training_json(tj, n_vocab=N_VOCAB)
    net = rh.build_reference_model(synthetic.make_random_checkpoint(cfg, 11), cfg=tj, n_vocab=N_VOCAB)
    path = rh.export_reference_onnx(tmp_path / "model.onnx", net, n_vocab=N_VOCAB)
    got = ow.config_from_onnx(str(path))
    assert got == cfg
    blob, man = weights.pack(ow.state_dict_from_onnx(str(path)), got, tc=False)
    b2, m2 = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 11)), cfg, tc=False)
    assert man == m2 and float(np.abs(blob - b2).max()) < 1e-6
