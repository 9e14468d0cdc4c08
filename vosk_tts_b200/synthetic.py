"""Seeded synthetic checkpoints in the reference's ``G_*.pth['model']`` layout.

No checkpoint (``vosk-model-tts-ru-0.9-multi`` / ``G_*.pth``) exists on the build or GPU
boxes, so benchmarks and parity tests run on random weights of the reference
architecture.  The tensor names and shapes follow
``SynthesizerTrn.state_dict()`` before weight-norm removal (SURVEY.md appendix B;
/root/reference/training/vits2/models.py:1508-1630): weight-normed convs appear as
``weight_g`` / ``weight_v`` pairs exactly as in a training checkpoint, so the loader's
folding path (``weights.fold_weight_norm``) is exercised.

Unlike the reference's own initialisation, the tensors that the reference zero-inits
(``flow.*.post`` models.py:371-372, ``ConvFlow.proj`` modules.py:361-362, LayerNorm beta)
are given non-zero values here, otherwise the coupling flow and the spline would be
identity maps and the parity tests would not exercise them.

The generator draws every tensor from its own ``torch.Generator`` (seeded from the
tensor name), so the values do not depend on enumeration order or torch's global RNG.
"""
import math
import zlib

import torch


def _spec(cfg):
    """Ordered list of (name, shape, kind, scale) for every tensor ``infer`` touches."""
    H = cfg["hidden_channels"]
    I = cfg["inter_channels"]
    Fc = cfg["filter_channels"]
    G = cfg["gin_channels"]
    nh = cfg["n_heads"]
    W = cfg["window_size"]
    k = cfg["kernel_size"]
    out = []

    def conv(name, co, ci, ks, wn=False, bias=True, gain=1.0, transposed=False):
        shape = (ci, co, ks) if transposed else (co, ci, ks)
        fan_in = ci * ks if not transposed else ci * ks / 4.0
        std = gain / math.sqrt(fan_in)
        if wn:
            out.append((name + ".weight_g", (shape[0], 1, 1), "wn_g", name + ".weight_v"))
            out.append((name + ".weight_v", shape, "normal", std))
        else:
            out.append((name + ".weight", shape, "normal", std))
        if bias:
            out.append((name + ".bias", (co,), "normal", 0.05))

    def ln(name, c):
        out.append((name + ".gamma", (c,), "gamma", 0.1))
        out.append((name + ".beta", (c,), "normal", 0.1))

    def encoder(prefix, hidden, filt, n_layers, ks):
        for i in range(n_layers):
            a = "%s.attn_layers.%d" % (prefix, i)
            dk = hidden // nh
            out.append((a + ".emb_rel_k", (1, 2 * W + 1, dk), "normal", dk ** -0.5))
            out.append((a + ".emb_rel_v", (1, 2 * W + 1, dk), "normal", dk ** -0.5))
            for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
                conv(a + "." + nm, hidden, hidden, 1, gain=1.0)
            ln("%s.norm_layers_1.%d" % (prefix, i), hidden)
            conv("%s.ffn_layers.%d.conv_1" % (prefix, i), filt, hidden, ks, gain=1.2)
            conv("%s.ffn_layers.%d.conv_2" % (prefix, i), hidden, filt, ks, gain=1.2)
            ln("%s.norm_layers_2.%d" % (prefix, i), hidden)

    def dds(prefix, c, ks, n_layers):
        for i in range(n_layers):
            out.append(("%s.convs_sep.%d.weight" % (prefix, i), (c, 1, ks), "normal", 0.4))
            out.append(("%s.convs_sep.%d.bias" % (prefix, i), (c,), "normal", 0.2))
            conv("%s.convs_1x1.%d" % (prefix, i), c, c, 1, gain=1.0)
            ln("%s.norms_1.%d" % (prefix, i), c)
            ln("%s.norms_2.%d" % (prefix, i), c)

    # --- speaker table, text encoder (models.py:283-326, attentions.py:13-65)
    if cfg["n_speakers"] > 1:
        out.append(("emb_g.weight", (cfg["n_speakers"], G), "normal", 1.0))
    out.append(("enc_p.emb.weight", (cfg["n_vocab"], H), "normal", H ** -0.5))
    encoder("enc_p.encoder", H, Fc, cfg["n_layers"], k)
    if cfg["use_spk_conditioned_encoder"] and G > 0:
        out.append(("enc_p.encoder.spk_emb_linear.weight", (H, G), "normal", 0.5 / math.sqrt(G)))
        out.append(("enc_p.encoder.spk_emb_linear.bias", (H,), "normal", 0.05))
    conv("enc_p.proj", 2 * I, H, 1, gain=0.6)

    # --- stochastic duration predictor (models.py:23-63); flows.1 is unused in reverse (:94-95)
    D = cfg["dp_filter_channels"]
    conv("dp.pre", D, H, 1)
    conv("dp.proj", D, D, 1)
    dds("dp.convs", D, cfg["dp_kernel_size"], 3)
    if G > 0:
        conv("dp.cond", D, G, 1, gain=0.5)
    out.append(("dp.flows.0.m", (2, 1), "ea_m", 0.2))
    out.append(("dp.flows.0.logs", (2, 1), "normal", 0.2))
    nb = cfg["dp_num_bins"]
    for f in range(cfg["dp_n_flows"]):
        p = "dp.flows.%d" % (2 * f + 1)
        conv(p + ".pre", D, 1, 1, gain=0.6)
        dds(p + ".convs", D, cfg["dp_kernel_size"], 3)
        out.append((p + ".proj.weight", (3 * nb - 1, D, 1), "spline_proj", nb))
        out.append((p + ".proj.bias", (3 * nb - 1,), "normal", 0.3))

    # --- flow (models.py:329-396 / 765-810, modules.py:111-184)
    fk = cfg["flow_kernel_size"]
    for f in range(cfg["flow_n_flows"]):
        p = "flow.flows.%d" % (2 * f)
        conv(p + ".pre", H, I // 2, 1)
        if cfg["use_transformer_flows"]:
            encoder(p + ".pre_transformer", H, H, 1, fk)
        for i in range(cfg["flow_wn_layers"]):
            conv("%s.enc.in_layers.%d" % (p, i), 2 * H, H, fk, wn=True, gain=1.0)
            rs = 2 * H if i < cfg["flow_wn_layers"] - 1 else H
            conv("%s.enc.res_skip_layers.%d" % (p, i), rs, H, 1, wn=True, gain=0.7)
        if G > 0:
            conv(p + ".enc.cond_layer", 2 * H * cfg["flow_wn_layers"], G, 1, wn=True, gain=0.5)
        conv(p + ".post", I // 2, H, 1, gain=0.35)

    # --- decoder (models.py:974-1063 / 845-898, modules.py:187-258)
    C0 = cfg["upsample_initial_channel"]
    if cfg["decoder"] in ("mb_istft", "ms_istft", "istft"):
        conv("dec.conv_pre", C0, I, 7, wn=True, gain=1.0)
    else:      # plain Generator: conv_pre is not weight-normed and a speaker projection is added to its output (models.py:851,869-875)
        conv("dec.conv_pre", C0, I, 7, wn=False, gain=1.0)
        if G > 0:
            conv("dec.cond", C0, G, 1, gain=0.5)
    ch = C0
    for i, (u, ku) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        conv("dec.ups.%d" % i, ch // 2, ch, ku, wn=True, gain=1.0, transposed=True)
        ch //= 2
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            rb = "dec.resblocks.%d" % (i * len(cfg["resblock_kernel_sizes"]) + j)
            if cfg["resblock"] == "1":
                for d in range(len(rd)):
                    conv("%s.convs1.%d" % (rb, d), ch, ch, rk, wn=True, gain=0.55)
                    conv("%s.convs2.%d" % (rb, d), ch, ch, rk, wn=True, gain=0.55)
            else:
                for d in range(len(rd)):
                    conv("%s.convs.%d" % (rb, d), ch, ch, rk, wn=True, gain=0.55)
    if cfg["decoder"] == "mb_istft":
        conv("dec.subband_conv_post", cfg["subbands"] * (cfg["gen_istft_n_fft"] + 2), ch, 7,
             wn=True, bias=False, gain=0.25)
    elif cfg["decoder"] == "ms_istft":      # models.py:1095 (bias!), :1107 learned 63-tap merge filter
        conv("dec.subband_conv_post", cfg["subbands"] * (cfg["gen_istft_n_fft"] + 2), ch, 7, wn=True, bias=True, gain=0.25)
        conv("dec.multistream_conv_post", 1, cfg["subbands"], 63, wn=True, bias=False, gain=1.0)
    elif cfg["decoder"] == "istft":         # models.py:928
        conv("dec.conv_post", cfg["gen_istft_n_fft"] + 2, ch, 7, wn=True, bias=False, gain=0.25)
    else:
        conv("dec.conv_post", 1, ch, 7, wn=False, bias=False, gain=0.5)     # plain Conv1d (models.py:868)
    return out


def _gen(name, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_random_checkpoint(cfg, seed=1234):
    """state_dict (CPU fp32) in checkpoint layout; deterministic for (cfg, seed)."""
    sd = {}
    spec = _spec(cfg)
    for name, shape, kind, arg in spec:
        if kind == "wn_g":
            continue
        g = _gen(name, seed)
        if kind == "normal":
            t = torch.randn(shape, generator=g) * arg
        elif kind == "ea_m":
            # shift log-durations up so that synthetic utterances average ~2 frames per token
            t = torch.randn(shape, generator=g) * arg - 0.9
        elif kind == "gamma":
            t = 1.0 + torch.randn(shape, generator=g) * arg
        elif kind == "spline_proj":
            nb = arg
            t = torch.randn(shape, generator=g) / math.sqrt(shape[1])
            t[: 2 * nb] *= 6.0   # widths / heights logits (divided by sqrt(filter) downstream)
            t[2 * nb:] *= 0.8    # derivative logits
        else:
            raise ValueError(kind)
        sd[name] = t.float().contiguous()
    for name, shape, kind, arg in spec:
        if kind != "wn_g":
            continue
        v = sd[arg]
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
        g = _gen(name, seed)
        sd[name] = (nrm * (1.0 + 0.15 * torch.randn(shape, generator=g))).float().contiguous()
    return sd


def param_names(cfg):
    return [s[0] for s in _spec(cfg)]
