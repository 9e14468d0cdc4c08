"""Checkpoint -> engine weights: weight-norm folding and the packed device blob.

Contract (SURVEY.md section 5 "Checkpoint / resume"; /root/reference/training/vits2/utils.py:18-50,
onnx_export.py:55,78-79): a ``G_*.pth`` holds ``{'model': state_dict, ...}`` with weight-normed convs
stored as ``weight_g`` / ``weight_v``; the exporter loads it and removes weight norm on ``dec`` and
``flow`` before tracing.  ``fold_weight_norm`` performs the same fold; ``pack`` (below) lays every
tensor out the way the CUDA kernels consume it.
"""
import numpy as np
import torch


def fold_weight_norm(sd):
    """w = g * v / ||v||, the norm taken over all dims but 0 (torch.nn.utils.weight_norm dim=0;
    for ConvTranspose1d dim 0 is the *input* channel, e.g. dec.ups.0.weight_g is [512,1,1])."""
    out = {}
    sd = {k: (torch.from_numpy(np.array(v)) if isinstance(v, np.ndarray) else v) for k, v in sd.items()}   # numpy (model.onnx) or torch
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            base = k[: -len("_v")]
            g = sd[base + "_g"]
            t = v.float()
            nrm = t.reshape(t.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (t.dim() - 1))
            out[base] = (t * (g.float() / nrm)).contiguous()
        elif k.endswith(".weight_g"):
            continue
        else:
            out[k] = v.float().contiguous() if torch.is_floating_point(v) else v
    return out


def load_checkpoint(path, allow_pickle=False):
    """Reads a reference ``G_*.pth`` (utils.py:18-21) and returns the folded state dict.

    The ``{'model': state_dict, 'iteration': int, 'optimizer': ..., 'learning_rate': float}`` layout is plain tensors
    and numbers, so the file is read with ``weights_only=True``: a crafted checkpoint in a model directory cannot run
    code at load time (the reference's inference package only ever opens ``model.onnx``).  ``allow_pickle=True`` is an
    explicit opt-in for legacy files that need full unpickling."""
    ck = torch.load(path, map_location="cpu", weights_only=not allow_pickle)
    sd = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
    return fold_weight_norm(sd)


# ----------------------------------------------------------------------------------------------
# Packing: folded state dict -> one fp32 blob + manifest, laid out the way csrc/ consumes it.
#   generic conv  <name>.w : [k][Cin][ldw]  (ldw = Cout rounded up to 4, zero padded), <name>.b : [ldw]
#   every tensor starts on a 256-byte boundary.
# ----------------------------------------------------------------------------------------------
def _kaiser(M, beta):
    n = np.arange(M)
    alpha = (M - 1) / 2.0
    return np.i0(beta * np.sqrt(np.clip(1 - ((n - alpha) / alpha) ** 2, 0, 1))) / np.i0(beta)


def istft_inverse_basis(n_fft, hop):
    """OnnxSTFT.__init__ (training/vits2/stft.py:191-214): pinv(scale*[Re;Im]FFT(I)[:n/2+1]).T * hann."""
    scale = n_fft / hop
    fb = np.fft.fft(np.eye(n_fft))
    cutoff = n_fft // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff]), np.imag(fb[:cutoff])])
    inv = np.linalg.pinv(scale * fb).T.astype(np.float32)          # FloatTensor(...) cast (:199-200)
    n = np.arange(n_fft)
    hann = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(np.float32)   # get_window('hann', fftbins=True).float()
    return (inv * hann[None, :]).astype(np.float32)                # [2*cutoff, n_fft]


def pqmf_synthesis_filter(subbands=4, taps=62, cutoff_ratio=0.15, beta=9.0):
    """PQMF.__init__ (training/vits2/pqmf.py:15-43,63-89): fp32 [subbands, taps+1]."""
    n = np.arange(taps + 1)
    omega_c = np.pi * cutoff_ratio
    with np.errstate(invalid="ignore", divide="ignore"):
        h_i = np.sin(omega_c * (n - 0.5 * taps)) / (np.pi * (n - 0.5 * taps))
    h_i[taps // 2] = np.cos(0) * cutoff_ratio
    h = h_i * _kaiser(taps + 1, beta)
    hs = np.zeros((subbands, taps + 1))
    for k in range(subbands):
        hs[k] = 2 * h * np.cos((2 * k + 1) * (np.pi / (2 * subbands)) * (n - ((taps - 1) / 2))
                               - (-1) ** k * np.pi / 4)
    return hs.astype(np.float32)


def to_bf16_bits(a):
    """float32 -> bf16 bit patterns (uint16), round-to-nearest-even like __float2bfloat16_rn."""
    x = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = (x + 0x7FFF + ((x >> 16) & 1)) >> 16
    return r.astype(np.uint16)


def from_bf16_bits(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


class _Packer:
    ALIGN = 64  # floats

    def __init__(self):
        self.chunks = []
        self.entries = []
        self.pos = 0

    def add(self, name, arr):
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32)).reshape(-1)
        pad = (-self.pos) % self.ALIGN
        if pad:
            self.chunks.append(np.zeros(pad, np.float32))
            self.pos += pad
        self.entries.append((name, self.pos, a.size))
        self.chunks.append(a)
        self.pos += a.size

    def conv(self, name, w, b=None, co_perm=None, ci_perm=None, need_w=True):
        """w: [Co, Ci, k] (Conv1d layout).  need_w=False: only the bias is packed (the conv runs on tcgen05 from its .th/.tl copy)."""
        w = np.asarray(w, np.float32)
        if co_perm is not None:
            w = w[co_perm]
            b = None if b is None else np.asarray(b, np.float32)[co_perm]
        if ci_perm is not None:
            w = w[:, ci_perm]
        co, ci, k = w.shape
        ldw = (co + 3) // 4 * 4
        wp = np.zeros((k, ci, ldw), np.float32)
        wp[:, :, :co] = np.transpose(w, (2, 1, 0))
        bp = np.zeros(ldw, np.float32)
        if b is not None:
            bp[:co] = np.asarray(b, np.float32)
        if need_w:
            self.add(name + ".w", wp)
        self.add(name + ".b", bp)

    def conv_tc(self, name, w, co_perm=None, ci_perm=None):
        """Split-bf16 copy of a conv weight for the tcgen05 path: <name>.th / <name>.tl = [k][Cout][Cin] bf16
        (K-major rows for TMA), hi = rne_bf16(w), lo = rne_bf16(w - hi); two bf16 per fp32 blob slot."""
        w = np.asarray(w, np.float32)
        if co_perm is not None:
            w = w[co_perm]
        if ci_perm is not None:
            w = w[:, ci_perm]
        wt = np.ascontiguousarray(np.transpose(w, (2, 0, 1)))          # [k][Cout][Cin]
        hi = to_bf16_bits(wt)
        lo = to_bf16_bits(wt - from_bf16_bits(hi))
        assert wt.size % 2 == 0
        self.add(name + ".th", hi.reshape(-1).view(np.float32))
        self.add(name + ".tl", lo.reshape(-1).view(np.float32))

    def conv_tc3(self, name, w, co_perm=None, ci_perm=None):
        """Exact 3-way split copy (precision mode 3): <name>.t3h / .t3m / .t3l with hi + mid + lo == w to the last fp32 bit."""
        w = np.asarray(w, np.float32)
        if co_perm is not None:
            w = w[co_perm]
        if ci_perm is not None:
            w = w[:, ci_perm]
        wt = np.ascontiguousarray(np.transpose(w, (2, 0, 1)))          # [k][Cout][Cin]
        hi = to_bf16_bits(wt)
        r1 = wt - from_bf16_bits(hi)
        mid = to_bf16_bits(r1)
        lo = to_bf16_bits(r1 - from_bf16_bits(mid))
        self.add(name + ".t3h", hi.reshape(-1).view(np.float32))
        self.add(name + ".t3m", mid.reshape(-1).view(np.float32))
        self.add(name + ".t3l", lo.reshape(-1).view(np.float32))

    def finish(self):
        blob = np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.float32)
        manifest = "".join("%s %d %d\n" % e for e in self.entries)
        return blob, manifest


def convt_phases(u, K):
    """Polyphase split of ConvTranspose1d(k=K, stride=u, padding=(K-u)//2): for phase r the taps
    (in increasing input position) are kernel columns j_m, and `pad` inputs lie left of t.
    out[u*t + r] = sum_m x[t - pad + m] * W[:, :, j_m]."""
    p = (K - u) // 2
    phases = []
    for r in range(u):
        d_min = -((r + p) // u)            # ceil(-(r+p)/u)
        d_max = (K - 1 - r - p) // u
        js = [r + p + u * (d_max - m) for m in range(d_max - d_min + 1)]
        assert all(0 <= j < K for j in js)
        phases.append((d_max, js))
    return phases


def tc_supported(cfg):
    """The tcgen05 conv path packs 64-channel K chunks: every conv it takes over must have Cin % 64 == 0."""
    n_ups = len(cfg["upsample_rates"])
    return (cfg["decoder"] in ("mb_istft", "ms_istft", "istft") and str(cfg["resblock"]) == "1" and cfg["hidden_channels"] % 64 == 0 and
            cfg["filter_channels"] % 64 == 0 and cfg["inter_channels"] % 64 == 0 and
            (cfg["upsample_initial_channel"] >> n_ups) % 64 == 0)


def pack(w, cfg, tc=True, precision=None):
    """w: folded state dict (reference names); returns (blob float32[n], manifest str).
    tc=True also packs split-bf16 copies of the convs for the tcgen05 path (precision modes 1 / 2).
    precision: None packs everything (a blob any engine mode can be created from); 0 / 1 / 2 leave out the tensors that
    mode never reads (mode 0: no split-bf16 copies at all; mode 1: none for the text encoder) -- what travels in the
    one-time NCCL weight broadcast of a multi-GPU job."""
    tc = tc and tc_supported(cfg) and precision != 0
    enc_tc = tc and precision in (None, 2)
    enc_tc3 = tc and precision in (None, 3)    # exact 3-way split copies of the text encoder's convs (mode 3)
    fw = not (tc and precision in (1, 2, 3))   # fp32 copies of the flow / decoder convs (the tcgen05 modes read only .th/.tl + bias)
    ew = not (tc and precision in (2, 3))      # ... of the text encoder's convs
    g = lambda k: w[k].detach().cpu().numpy() if hasattr(w[k], "detach") else np.asarray(w[k])
    H, I, G = cfg["hidden_channels"], cfg["inter_channels"], cfg["gin_channels"]
    D = cfg["dp_filter_channels"]
    P = _Packer()

    def ln(dst, src):
        P.add(dst + ".g", g(src + ".gamma"))
        P.add(dst + ".b", g(src + ".beta"))

    def enc_layer(dst, src, i, with_tc=False, need_w=True, with_tc3=False):
        a = "%s.attn_layers.%d" % (src, i)
        wq = np.concatenate([g(a + ".conv_q.weight"), g(a + ".conv_k.weight"), g(a + ".conv_v.weight")], 0)
        bq = np.concatenate([g(a + ".conv_q.bias"), g(a + ".conv_k.bias"), g(a + ".conv_v.bias")], 0)
        P.conv(dst + ".qkv", wq, bq, need_w=need_w)
        P.conv(dst + ".o", g(a + ".conv_o.weight"), g(a + ".conv_o.bias"), need_w=need_w)
        if with_tc:
            f_ = "%s.ffn_layers.%d" % (src, i)
            P.conv_tc(dst + ".qkv", wq)
            P.conv_tc(dst + ".o", g(a + ".conv_o.weight"))
            P.conv_tc(dst + ".ffn1", g(f_ + ".conv_1.weight"))
            P.conv_tc(dst + ".ffn2", g(f_ + ".conv_2.weight"))
        if with_tc3:
            f_ = "%s.ffn_layers.%d" % (src, i)
            P.conv_tc3(dst + ".qkv", wq)
            P.conv_tc3(dst + ".o", g(a + ".conv_o.weight"))
            P.conv_tc3(dst + ".ffn1", g(f_ + ".conv_1.weight"))
            P.conv_tc3(dst + ".ffn2", g(f_ + ".conv_2.weight"))
        P.add(dst + ".relk", g(a + ".emb_rel_k")[0])
        P.add(dst + ".relv", g(a + ".emb_rel_v")[0])
        if with_tc:
            # the same tables as split-bf16 [16 offsets][128 channels] tiles (zero padded) for the tcgen05 attention:
            # Ek is a K-major B operand of Q Ek^T, Ev an MN-major B operand of P_band Ev (csrc/attn_tc.cuh)
            for nm, src_t in ((".rk", g(a + ".emb_rel_k")[0]), (".rv", g(a + ".emb_rel_v")[0])):
                nrel, dk = src_t.shape
                if nrel <= 16 and dk <= 128:
                    t = np.zeros((16, 128), np.float32)
                    t[:nrel, :dk] = src_t
                    hi = to_bf16_bits(t)
                    lo = to_bf16_bits(t - from_bf16_bits(hi))
                    P.add(dst + nm + "h", hi.reshape(-1).view(np.float32))
                    P.add(dst + nm + "l", lo.reshape(-1).view(np.float32))
        ln(dst + ".ln1", "%s.norm_layers_1.%d" % (src, i))
        f = "%s.ffn_layers.%d" % (src, i)
        P.conv(dst + ".ffn1", g(f + ".conv_1.weight"), g(f + ".conv_1.bias"), need_w=need_w)
        P.conv(dst + ".ffn2", g(f + ".conv_2.weight"), g(f + ".conv_2.bias"), need_w=need_w)
        ln(dst + ".ln2", "%s.norm_layers_2.%d" % (src, i))

    def dds(dst, src, n_layers=3):
        for i in range(n_layers):
            P.add("%s.%d.sep_w" % (dst, i), np.transpose(g("%s.convs_sep.%d.weight" % (src, i))[:, 0, :], (1, 0)))
            P.add("%s.%d.sep_b" % (dst, i), g("%s.convs_sep.%d.bias" % (src, i)))
            ln("%s.%d.ln1" % (dst, i), "%s.norms_1.%d" % (src, i))
            P.conv("%s.%d.pw" % (dst, i), g("%s.convs_1x1.%d.weight" % (src, i)), g("%s.convs_1x1.%d.bias" % (src, i)))
            ln("%s.%d.ln2" % (dst, i), "%s.norms_2.%d" % (src, i))

    # ---- speaker table + all per-utterance conditioning projections as one matrix
    has_g = cfg["n_speakers"] > 0 and G > 0
    if has_g:
        P.add("emb_g", g("emb_g.weight"))
        rows_w, rows_b = [], []
        if cfg["use_spk_conditioned_encoder"]:
            rows_w.append(g("enc_p.encoder.spk_emb_linear.weight"))
            rows_b.append(g("enc_p.encoder.spk_emb_linear.bias"))
        rows_w.append(g("dp.cond.weight")[:, :, 0])
        rows_b.append(g("dp.cond.bias"))
        nl = cfg["flow_wn_layers"]
        il = np.arange(2 * H).reshape(2, H).T.reshape(-1)     # gate interleave: [t0,s0,t1,s1,...]
        for f in range(cfg["flow_n_flows"]):
            cw = g("flow.flows.%d.enc.cond_layer.weight" % (2 * f))[:, :, 0]
            cb = g("flow.flows.%d.enc.cond_layer.bias" % (2 * f))
            for i in range(nl):
                rows_w.append(cw[i * 2 * H:(i + 1) * 2 * H][il])
                rows_b.append(cb[i * 2 * H:(i + 1) * 2 * H][il])
        if cfg["decoder"] == "hifigan" and "dec.cond.weight" in w:
            rows_w.append(g("dec.cond.weight")[:, :, 0])       # Generator's speaker projection (models.py:869-875)
            rows_b.append(g("dec.cond.bias"))
        P.add("cond.w", np.concatenate(rows_w, 0))
        P.add("cond.b", np.concatenate(rows_b, 0))

    # ---- text encoder
    P.add("enc.emb", g("enc_p.emb.weight"))
    for i in range(cfg["n_layers"]):
        enc_layer("enc.%d" % i, "enc_p.encoder", i, with_tc=enc_tc, need_w=ew, with_tc3=enc_tc3)
    P.conv("enc.proj", g("enc_p.proj.weight"), g("enc_p.proj.bias"), need_w=ew)
    if enc_tc:
        P.conv_tc("enc.proj", g("enc_p.proj.weight"))
    if enc_tc3:
        P.conv_tc3("enc.proj", g("enc_p.proj.weight"))

    # ---- stochastic duration predictor
    P.conv("dp.pre", g("dp.pre.weight"), g("dp.pre.bias"))
    P.conv("dp.proj", g("dp.proj.weight"), g("dp.proj.bias"))
    dds("dp.convs", "dp.convs")
    for n in range(2, cfg["dp_n_flows"] + 1):
        src = "dp.flows.%d" % (2 * n - 1)
        P.add("dp.cf%d.pre_w" % n, g(src + ".pre.weight")[:, 0, 0])
        P.add("dp.cf%d.pre_b" % n, g(src + ".pre.bias"))
        dds("dp.cf%d.convs" % n, src + ".convs")
        P.conv("dp.cf%d.proj" % n, g(src + ".proj.weight"), g(src + ".proj.bias"))
    P.add("dp.ea", np.concatenate([g("dp.flows.0.m").reshape(-1), g("dp.flows.0.logs").reshape(-1)]))

    # ---- flow (reverse); channel flips are folded into the pre/post weights (see csrc/engine.cu)
    nf = cfg["flow_n_flows"]
    half = I // 2
    rev = np.arange(half)[::-1].copy()
    for f in range(nf):
        src = "flow.flows.%d" % (2 * f)
        dst = "flow.%d" % f
        flipped = ((nf - f) % 2) == 1
        P.conv(dst + ".pre", g(src + ".pre.weight"), g(src + ".pre.bias"), ci_perm=rev if flipped else None)
        if cfg["use_transformer_flows"]:
            enc_layer(dst + ".tr", src + ".pre_transformer", 0, with_tc=tc, need_w=fw)
        nl = cfg["flow_wn_layers"]
        il = np.arange(2 * H).reshape(2, H).T.reshape(-1)
        for i in range(nl):
            P.conv("%s.in%d" % (dst, i), g("%s.enc.in_layers.%d.weight" % (src, i)),
                   g("%s.enc.in_layers.%d.bias" % (src, i)), co_perm=il, need_w=fw)
            rw, rb = g("%s.enc.res_skip_layers.%d.weight" % (src, i)), g("%s.enc.res_skip_layers.%d.bias" % (src, i))
            if tc:
                P.conv_tc("%s.in%d" % (dst, i), g("%s.enc.in_layers.%d.weight" % (src, i)), co_perm=il)
            if i < nl - 1:
                P.conv("%s.rsx%d" % (dst, i), rw[:H], rb[:H], need_w=fw)
                P.conv("%s.rss%d" % (dst, i), rw[H:], rb[H:], need_w=fw)
                if tc:
                    P.conv_tc("%s.rsx%d" % (dst, i), rw[:H])
                    P.conv_tc("%s.rss%d" % (dst, i), rw[H:])
            else:
                P.conv("%s.rss%d" % (dst, i), rw, rb, need_w=fw)
                if tc:
                    P.conv_tc("%s.rss%d" % (dst, i), rw)
        P.conv(dst + ".post", g(src + ".post.weight"), g(src + ".post.bias"), co_perm=rev if flipped else None, need_w=fw)
        if tc:
            P.conv_tc(dst + ".post", g(src + ".post.weight"), co_perm=rev if flipped else None)

    # ---- decoder
    pre_w = g("dec.conv_pre.weight")
    if nf % 2 == 1:   # odd number of flips leaves the latent channel-reversed: fold into conv_pre
        pre_w = pre_w[:, ::-1].copy()
    P.conv("dec.pre", pre_w, g("dec.conv_pre.bias"), need_w=fw)
    if tc:
        P.conv_tc("dec.pre", pre_w)
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, ku) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        wt = g("dec.ups.%d.weight" % i)                      # [Cin, Cout, K]
        bt = g("dec.ups.%d.bias" % i)
        for r, (pad, js) in enumerate(convt_phases(u, ku)):
            wr = np.stack([wt[:, :, j] for j in js], axis=-1)   # [Cin, Cout, ntaps]
            P.conv("dec.up%d.p%d" % (i, r), np.transpose(wr, (1, 0, 2)), bt, need_w=fw)
            if tc:
                P.conv_tc("dec.up%d.p%d" % (i, r), np.transpose(wr, (1, 0, 2)))
        for j in range(nk):
            n = i * nk + j
            nd = len(cfg["resblock_dilation_sizes"][j])
            for d in range(nd):
                if cfg["resblock"] == "1":
                    P.conv("dec.rb%d.c1.%d" % (n, d), g("dec.resblocks.%d.convs1.%d.weight" % (n, d)), g("dec.resblocks.%d.convs1.%d.bias" % (n, d)), need_w=fw)
                    P.conv("dec.rb%d.c2.%d" % (n, d), g("dec.resblocks.%d.convs2.%d.weight" % (n, d)), g("dec.resblocks.%d.convs2.%d.bias" % (n, d)), need_w=fw)
                    if tc:
                        P.conv_tc("dec.rb%d.c1.%d" % (n, d), g("dec.resblocks.%d.convs1.%d.weight" % (n, d)))
                        P.conv_tc("dec.rb%d.c2.%d" % (n, d), g("dec.resblocks.%d.convs2.%d.weight" % (n, d)))
                else:
                    P.conv("dec.rb%d.c.%d" % (n, d), g("dec.resblocks.%d.convs.%d.weight" % (n, d)), g("dec.resblocks.%d.convs.%d.bias" % (n, d)))
    if cfg["decoder"] in ("mb_istft", "ms_istft", "istft"):
        # All three end in conv_post -> exp / pi*sin -> inverse STFT -> zero-stuffing by `subbands` -> a 63-tap filter per band
        # (zero padding 31).  Only the filter differs: the fixed PQMF synthesis bank (pqmf.py:63-89), the learned
        # multistream_conv_post (models.py:1107), or -- one band, nothing after the iSTFT (models.py:962-965) -- a unit impulse.
        post = "dec.conv_post" if cfg["decoder"] == "istft" else "dec.subband_conv_post"
        post_b = g(post + ".bias") if (post + ".bias") in w else None          # only the multistream decoder has one (:1095)
        P.conv("dec.post", g(post + ".weight"), post_b, need_w=fw)
        if tc:
            P.conv_tc("dec.post", g(post + ".weight"))
        P.add("dec.istft", istft_inverse_basis(cfg["gen_istft_n_fft"], cfg["gen_istft_hop_size"]))
        if cfg["decoder"] == "mb_istft":
            bank = pqmf_synthesis_filter(cfg["subbands"])
        elif cfg["decoder"] == "ms_istft":
            bank = g("dec.multistream_conv_post.weight")[0]
            assert bank.shape == (cfg["subbands"], 63), bank.shape
        else:
            bank = np.zeros((1, 63), np.float32)
            bank[0, 31] = 1.0
        P.add("dec.pqmf", bank)
    else:
        P.conv("dec.post", g("dec.conv_post.weight"), None)
    return P.finish()
