"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference VITS2 inference path.

This module is the *oracle* for the parity tests and the timed CPU baseline of
``bench.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs may import it; the product path
(``vosk_tts_b200``) never does, and has no CPU fallback.

It restates, as plain functions over a dict of folded fp32 tensors, what
``SynthesizerTrn.infer`` (/root/reference/training/vits2/models.py:1679-1704) computes --
the graph that the deployed ``model.onnx`` is a trace of (onnx_export.py:47-104).  It uses
the same ATen CPU kernels (conv1d / matmul / layer_norm ...) in the same order as the
reference modules, so it is also an honest stand-in for timing the reference's CPU path
(onnxruntime and model.onnx are absent from this image).

Pinning: ``tests/test_oracle_vs_reference.py`` runs this restatement against the
unmodified reference modules (imported from /root/reference, build container only) and
``tests/golden/*.npz`` hold outputs of the *reference itself* on seeded inputs
(``oracle/make_golden.py``).  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so those reference-generated fixtures are the pin.

All tensors are [B, C, T] (the reference's layout), fp32.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules.py:17


# --------------------------------------------------------------------------- helpers
def sequence_mask(lengths, max_len=None):
    """commons.py:121-125."""
    if max_len is None:
        max_len = int(lengths.max())
    pos = torch.arange(max_len, dtype=lengths.dtype)
    return pos[None, :] < lengths[:, None]


def layer_norm_c(x, gamma, beta, eps=1e-5):
    """modules.py:29-32: LayerNorm over the channel dim of [B,C,T]."""
    y = F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps)
    return y.transpose(1, -1)


def conv(x, w, name, **kw):
    return F.conv1d(x, w[name + ".weight"], w.get(name + ".bias"), **kw)


# --------------------------------------------------------------------------- attention
def rel_attention(x, attn_mask, w, p, n_heads, window):
    """attentions.py:155-196 (MultiHeadAttention.forward/attention), self-attention only.

    The reference realises the windowed relative-position terms with pad/reshape "skew"
    tricks (:216-260); here the same quantities are gathered directly:
    score[i,j] += q_i . Ek[j-i+W] for |j-i| <= W, out_i += sum_j p[i,j] Ev[j-i+W]."""
    B, C, T = x.shape
    dk = C // n_heads
    q = conv(x, w, p + ".conv_q").view(B, n_heads, dk, T).transpose(2, 3)
    k = conv(x, w, p + ".conv_k").view(B, n_heads, dk, T).transpose(2, 3)
    v = conv(x, w, p + ".conv_v").view(B, n_heads, dk, T).transpose(2, 3)
    qs = q / math.sqrt(dk)
    scores = torch.matmul(qs, k.transpose(-2, -1))
    ii = torch.arange(T)
    rel = ii[None, :] - ii[:, None] + window            # j - i + W
    band = (rel >= 0) & (rel <= 2 * window)
    relc = rel.clamp(0, 2 * window)
    ek = w[p + ".emb_rel_k"]                             # [1, 2W+1, dk]
    rel_logits = torch.matmul(qs, ek.unsqueeze(0).transpose(-2, -1))   # [B,H,T,2W+1]
    local = torch.gather(rel_logits, -1, relc.expand(B, n_heads, T, T))
    scores = scores + local * band
    scores = scores.masked_fill(attn_mask == 0, -1e4)   # :183 (fill value is -1e4, not -inf)
    pr = F.softmax(scores, dim=-1)
    out = torch.matmul(pr, v)
    # relative weights r[i,m] = p[i, i+m-W] (zero outside the sequence)
    jj = ii[:, None] + torch.arange(2 * window + 1)[None, :] - window   # [T, 2W+1]
    ok = (jj >= 0) & (jj < T)
    rw = torch.gather(pr, -1, jj.clamp(0, T - 1).expand(B, n_heads, T, 2 * window + 1)) * ok
    out = out + torch.matmul(rw, w[p + ".emb_rel_v"].unsqueeze(0))
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return conv(out, w, p + ".conv_o")


def ffn(x, x_mask, w, p, ks):
    """attentions.py:294-320: conv -> ReLU -> conv with 'same' padding, masks."""
    pl, pr = (ks - 1) // 2, ks // 2
    y = conv(F.pad(x * x_mask, (pl, pr)), w, p + ".conv_1")
    y = torch.relu(y)
    y = conv(F.pad(y * x_mask, (pl, pr)), w, p + ".conv_2")
    return y * x_mask


def encoder(x, x_mask, w, p, n_layers, n_heads, ks, window, g=None, cond_layer_idx=None):
    """attentions.py:48-65."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    for i in range(n_layers):
        if g is not None and i == cond_layer_idx:
            gl = F.linear(g.transpose(1, 2), w[p + ".spk_emb_linear.weight"],
                          w[p + ".spk_emb_linear.bias"]).transpose(1, 2)
            x = (x + gl) * x_mask
        y = rel_attention(x, attn_mask, w, "%s.attn_layers.%d" % (p, i), n_heads, window)
        x = layer_norm_c(x + y, w["%s.norm_layers_1.%d.gamma" % (p, i)], w["%s.norm_layers_1.%d.beta" % (p, i)])
        y = ffn(x, x_mask, w, "%s.ffn_layers.%d" % (p, i), ks)
        x = layer_norm_c(x + y, w["%s.norm_layers_2.%d.gamma" % (p, i)], w["%s.norm_layers_2.%d.beta" % (p, i)])
    return x * x_mask


def text_encoder(tokens, lengths, g, w, cfg):
    """models.py:317-326."""
    H = cfg["hidden_channels"]
    x = F.embedding(tokens, w["enc_p.emb.weight"]) * math.sqrt(H)
    x = x.transpose(1, -1)
    x_mask = sequence_mask(lengths, x.shape[2]).unsqueeze(1).to(x.dtype)
    use_g = g if (cfg["use_spk_conditioned_encoder"] and cfg["gin_channels"] > 0) else None
    x = encoder(x * x_mask, x_mask, w, "enc_p.encoder", cfg["n_layers"], cfg["n_heads"],
                cfg["kernel_size"], cfg["window_size"], g=use_g, cond_layer_idx=cfg["cond_layer_idx"])
    stats = conv(x, w, "enc_p.proj") * x_mask
    m, logs = torch.split(stats, cfg["inter_channels"], dim=1)
    return x, m, logs, x_mask


# --------------------------------------------------------------------------- duration predictor
def dds_conv(x, x_mask, w, p, ks, n_layers, g=None):
    """modules.py:96-108."""
    if g is not None:
        x = x + g
    for i in range(n_layers):
        dil = ks ** i
        pad = (ks * dil - dil) // 2
        C = x.shape[1]
        y = F.conv1d(x * x_mask, w["%s.convs_sep.%d.weight" % (p, i)], w["%s.convs_sep.%d.bias" % (p, i)],
                     groups=C, dilation=dil, padding=pad)
        y = layer_norm_c(y, w["%s.norms_1.%d.gamma" % (p, i)], w["%s.norms_1.%d.beta" % (p, i)])
        y = F.gelu(y)
        y = conv(y, w, "%s.convs_1x1.%d" % (p, i))
        y = layer_norm_c(y, w["%s.norms_2.%d.gamma" % (p, i)], w["%s.norms_2.%d.beta" % (p, i)])
        y = F.gelu(y)
        x = x + y
    return x * x_mask


def rq_spline_inverse(x, uw, uh, ud, bound=5.0, min_w=1e-3, min_h=1e-3, min_d=1e-3):
    """transforms.py:55-193, inverse branch with linear tails.

    x: [...]; uw, uh: [..., nb]; ud: [..., nb-1].  Values outside [-bound, bound] pass through."""
    nb = uw.shape[-1]
    inside = (x >= -bound) & (x <= bound)
    const = float(np.log(np.exp(1 - min_d) - 1))                      # :73
    ud = F.pad(ud, (1, 1))
    ud[..., 0] = const
    ud[..., -1] = const

    def knots(u, min_b):
        b = F.softmax(u, dim=-1)
        b = min_b + (1 - min_b * nb) * b
        cb = F.pad(torch.cumsum(b, dim=-1), (1, 0), value=0.0)
        cb = 2 * bound * cb + (-bound)
        cb[..., 0] = -bound
        cb[..., -1] = bound
        return cb, cb[..., 1:] - cb[..., :-1]

    cw, widths = knots(uw, min_w)
    chh, heights = knots(uh, min_h)
    deriv = min_d + F.softplus(ud)
    xin = torch.where(inside, x, torch.zeros_like(x))                  # any in-domain value for masked-out lanes
    loc = chh.clone()
    loc[..., -1] += 1e-6                                               # :48
    bin_idx = (torch.sum(xin[..., None] >= loc, dim=-1) - 1)[..., None]
    g1 = lambda t: t.gather(-1, bin_idx)[..., 0]
    in_cw, in_w, in_ch, in_h = g1(cw), g1(widths), g1(chh), g1(heights)
    delta = heights / widths
    in_delta = g1(delta)
    d0 = g1(deriv)
    d1 = deriv[..., 1:].gather(-1, bin_idx)[..., 0]
    t = d0 + d1 - 2 * in_delta
    a = (xin - in_ch) * t + in_h * (in_delta - d0)
    b = in_h * d0 - (xin - in_ch) * t
    c = -in_delta * (xin - in_ch)
    disc = b.pow(2) - 4 * a * c
    root = (2 * c) / (-b - torch.sqrt(disc))
    y = root * in_w + in_cw
    return torch.where(inside, y, x)


def conv_flow_reverse(z, x_mask, w, p, g, cfg):
    """modules.py:365-392 with reverse=True."""
    D = cfg["dp_filter_channels"]
    nb = cfg["dp_num_bins"]
    x0, x1 = z[:, :1], z[:, 1:]
    h = conv(x0, w, p + ".pre")
    h = dds_conv(h, x_mask, w, p + ".convs", cfg["dp_kernel_size"], 3, g=g)
    h = conv(h, w, p + ".proj") * x_mask
    B, _, T = x0.shape
    h = h.reshape(B, 1, -1, T).permute(0, 1, 3, 2)
    uw = h[..., :nb] / math.sqrt(D)
    uh = h[..., nb:2 * nb] / math.sqrt(D)
    ud = h[..., 2 * nb:]
    x1 = rq_spline_inverse(x1, uw, uh, ud, bound=cfg["dp_tail_bound"])
    return torch.cat([x0, x1], 1) * x_mask


def sdp_reverse(x, x_mask, g, eps_dp, noise_scale_w, w, cfg):
    """models.py:56-63 and :93-101 (reverse=True).  eps_dp stands in for torch.randn(B,2,T) (:96)."""
    x = conv(x, w, "dp.pre")
    if g is not None:
        x = x + conv(g, w, "dp.cond")
    x = dds_conv(x, x_mask, w, "dp.convs", cfg["dp_kernel_size"], 3)
    x = conv(x, w, "dp.proj") * x_mask
    z = eps_dp * noise_scale_w
    nf = cfg["dp_n_flows"]
    # reversed(flows) with the first ConvFlow ("useless vflow", :94-95) dropped:
    # Flip, CF_nf, Flip, ..., CF_2, Flip, ElementwiseAffine
    for f in range(nf, 1, -1):
        z = torch.flip(z, [1])
        z = conv_flow_reverse(z, x_mask, w, "dp.flows.%d" % (2 * f - 1), x, cfg)
    z = torch.flip(z, [1])
    z = (z - w["dp.flows.0.m"]) * torch.exp(-w["dp.flows.0.logs"]) * x_mask     # modules.py:296
    return z[:, :1]


# --------------------------------------------------------------------------- length regulator
def durations(logw, x_mask, length_scale):
    """models.py:1689-1691."""
    wdur = torch.exp(logw) * x_mask * length_scale
    w_ceil = torch.ceil(wdur)
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    return w_ceil, y_lengths


def frame_to_token(w_ceil, y_lengths):
    """commons.generate_path (commons.py:128-143) as an index map: frame j of utterance b takes
    token idx[b,j] = #{i : cumsum_i <= j}.  Returns int64 [B, max T_y] (-1 beyond y_lengths)."""
    cum = torch.cumsum(w_ceil[:, 0, :], -1)                     # [B,T_x]
    Ty = int(y_lengths.max())
    j = torch.arange(Ty, dtype=cum.dtype)
    idx = (cum[:, None, :] <= j[None, :, None]).sum(-1)          # [B,Ty]
    valid = j[None, :] < y_lengths[:, None].to(cum.dtype)
    return torch.where(valid, idx, torch.full_like(idx, -1))


def path_from_index(idx, T_x):
    """Dense 0/1 attn [B,1,T_y,T_x] equivalent to commons.generate_path(...) * mask."""
    B, Ty = idx.shape
    attn = torch.zeros(B, Ty, T_x)
    ok = idx >= 0
    safe = idx.clamp(0, T_x - 1)
    attn.scatter_(2, safe[..., None], ok[..., None].float())
    # a frame whose token index runs past T_x (cannot happen when y_len = sum w_ceil) stays zero
    attn = attn * (idx < T_x)[..., None]
    return attn.unsqueeze(1)


# --------------------------------------------------------------------------- flow
def wn(x, x_mask, g, w, p, cfg):
    """modules.py:148-176 (weight norm already folded)."""
    H = cfg["hidden_channels"]
    ks = cfg["flow_kernel_size"]
    nl = cfg["flow_wn_layers"]
    out = torch.zeros_like(x)
    gc = conv(g, w, p + ".cond_layer") if g is not None else None
    for i in range(nl):
        dil = cfg["flow_dilation_rate"] ** i
        pad = int((ks * dil - dil) / 2)
        x_in = conv(x, w, "%s.in_layers.%d" % (p, i), dilation=dil, padding=pad)
        if gc is not None:
            x_in = x_in + gc[:, i * 2 * H:(i + 1) * 2 * H, :]
        acts = torch.tanh(x_in[:, :H]) * torch.sigmoid(x_in[:, H:])       # commons.py:100-107
        rs = conv(acts, w, "%s.res_skip_layers.%d" % (p, i))
        if i < nl - 1:
            x = (x + rs[:, :H]) * x_mask
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out * x_mask


def coupling_reverse(x, x_mask, g, w, p, cfg):
    """models.py:374-393 (ResidualCouplingTransformersLayer2, mean_only) /
    modules.py:326-345 (ResidualCouplingLayer, mean_only) with reverse=True."""
    half = cfg["inter_channels"] // 2
    x0, x1 = x[:, :half], x[:, half:]
    h = conv(x0, w, p + ".pre") * x_mask
    if cfg["use_transformer_flows"]:
        h = h + encoder(h * x_mask, x_mask, w, p + ".pre_transformer", 1, 2,
                        cfg["flow_kernel_size"], cfg["window_size"])
    h = wn(h, x_mask, g, w, p + ".enc", cfg)
    m = conv(h, w, p + ".post") * x_mask
    x1 = (x1 - m) * x_mask
    return torch.cat([x0, x1], 1)


def flow_reverse(z, y_mask, g, w, cfg):
    """models.py:750-757: reversed([L1, Flip, ..., Ln, Flip])."""
    for f in range(cfg["flow_n_flows"] - 1, -1, -1):
        z = torch.flip(z, [1])
        z = coupling_reverse(z, y_mask, g, w, "flow.flows.%d" % (2 * f), cfg)
    return z


# --------------------------------------------------------------------------- decoder
def resblock1(x, w, p, ks, dils):
    """modules.py:210-225 with x_mask=None."""
    for i, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = conv(xt, w, "%s.convs1.%d" % (p, i), dilation=d, padding=(ks * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = conv(xt, w, "%s.convs2.%d" % (p, i), padding=(ks - 1) // 2)
        x = xt + x
    return x


def resblock2(x, w, p, ks, dils):
    """modules.py:245-254 with x_mask=None."""
    for i, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = conv(xt, w, "%s.convs.%d" % (p, i), dilation=d, padding=(ks * d - d) // 2)
        x = xt + x
    return x


def istft_inverse_basis(n_fft, hop):
    """stft.py:191-214 (OnnxSTFT.__init__): pinv(scale * [Re;Im] FFT(I)[:n/2+1]).T * hann(periodic)."""
    scale = n_fft / hop
    fb = np.fft.fft(np.eye(n_fft))
    cutoff = n_fft // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff]), np.imag(fb[:cutoff])])
    inv = np.linalg.pinv(scale * fb).T                                    # [2*cutoff, n_fft]
    n = np.arange(n_fft)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)                    # scipy get_window('hann', fftbins=True)
    basis = torch.FloatTensor(inv[:, None, :])
    basis = basis * torch.from_numpy(hann).float()
    return basis.float()                                                  # [18,1,16]


def kaiser_window(M, beta):
    n = np.arange(M)
    alpha = (M - 1) / 2.0
    return np.i0(beta * np.sqrt(np.clip(1 - ((n - alpha) / alpha) ** 2, 0, 1))) / np.i0(beta)


def pqmf_synthesis_filter(subbands=4, taps=62, cutoff_ratio=0.15, beta=9.0):
    """pqmf.py:15-43 and :63-89: cosine-modulated synthesis bank, fp32 [1, subbands, taps+1]."""
    n = np.arange(taps + 1)
    omega_c = np.pi * cutoff_ratio
    with np.errstate(invalid="ignore", divide="ignore"):
        h_i = np.sin(omega_c * (n - 0.5 * taps)) / (np.pi * (n - 0.5 * taps))
    h_i[taps // 2] = np.cos(0) * cutoff_ratio
    h = h_i * kaiser_window(taps + 1, beta)
    hs = np.zeros((subbands, taps + 1))
    for k in range(subbands):
        hs[k] = 2 * h * np.cos((2 * k + 1) * (np.pi / (2 * subbands)) * (n - ((taps - 1) / 2))
                               - (-1) ** k * np.pi / 4)
    return torch.from_numpy(hs).float().unsqueeze(0)


def decoder_trunk(z, w, cfg, g=None):
    """conv_pre + upsample/MRF stages: models.py:1024-1036 (same trunk as Generator :873-885, which also adds cond(g))."""
    x = conv(z, w, "dec.conv_pre", padding=3)
    if g is not None and "dec.cond.weight" in w:
        x = x + conv(g, w, "dec.cond")                     # Generator.forward models.py:874-875
    nk = len(cfg["resblock_kernel_sizes"])
    rb = resblock1 if cfg["resblock"] == "1" else resblock2
    for i, (u, ku) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, w["dec.ups.%d.weight" % i], w["dec.ups.%d.bias" % i],
                               stride=u, padding=(ku - u) // 2)
        xs = None
        for j in range(nk):
            r = rb(x, w, "dec.resblocks.%d" % (i * nk + j), cfg["resblock_kernel_sizes"][j],
                   cfg["resblock_dilation_sizes"][j])
            xs = r if xs is None else xs + r
        x = xs / nk
    return x


def decoder_mb_istft(z, w, cfg):
    """The three inverse-STFT decoders, all with OnnxSTFT (is_onnx=True):
    mb_istft  Multiband_iSTFT_Generator.forward models.py:1016-1054 + PQMF.synthesis pqmf.py:105-116;
    ms_istft  Multistream_iSTFT_Generator.forward models.py:1117-1158: same, conv_post has a bias and the fixed PQMF
              synthesis bank is replaced by the learned 63-tap ``multistream_conv_post`` (zero padding 31, no bias);
    istft     iSTFT_Generator.forward models.py:947-965: one band, ``conv_post``, no filter bank."""
    kind = cfg["decoder"]
    sb = 1 if kind == "istft" else cfg["subbands"]
    nfft, hop = cfg["gen_istft_n_fft"], cfg["gen_istft_hop_size"]
    x = decoder_trunk(z, w, cfg)
    x = F.leaky_relu(x)                                   # default slope 0.01 (:1038)
    x = F.pad(x, (1, 0), mode="reflect")
    x = conv(x, w, "dec.conv_post" if kind == "istft" else "dec.subband_conv_post", padding=3)
    B, _, L = x.shape
    x = x.reshape(B, sb, x.shape[1] // sb, L)
    nb = nfft // 2 + 1
    spec = torch.exp(x[:, :, :nb, :]).reshape(B * sb, nb, L)
    phase = (math.pi * torch.sin(x[:, :, nb:, :])).reshape(B * sb, nb, L)
    rec = torch.cat([spec * torch.cos(phase), spec * torch.sin(phase)], dim=1)
    y = F.conv_transpose1d(rec, istft_inverse_basis(nfft, hop), stride=hop)
    y = y * (float(nfft) / hop)                           # stft.py:256
    y = y[:, :, nfft // 2:]
    y = y[:, :, :-(nfft // 2)]
    y_mb = y.reshape(B, sb, y.shape[-1])
    if kind == "istft":
        return y_mb, None
    updown = torch.zeros(sb, sb, sb)
    for k in range(sb):
        updown[k, k, 0] = 1.0
    up = F.conv_transpose1d(y_mb, updown * sb, stride=sb)
    if kind == "ms_istft":
        wav = F.conv1d(up, w["dec.multistream_conv_post.weight"], None, padding=31)      # get_padding(63, 1) = 31
        return wav, up                                    # the reference returns the zero-stuffed bands as y_mb_hat
    wav = F.conv1d(F.pad(up, (31, 31)), pqmf_synthesis_filter(sb))
    return wav, y_mb


def decoder_hifigan(z, w, cfg, g=None):
    """models.py:872-891 (plain HiFi-GAN Generator)."""
    x = decoder_trunk(z, w, cfg, g)
    x = F.leaky_relu(x)
    x = conv(x, w, "dec.conv_post", padding=3)
    return torch.tanh(x), None


# --------------------------------------------------------------------------- the whole path
def infer(w, cfg, tokens, lengths, sid, scales, eps_dp, eps_z=None, return_all=False, decode=True):
    """SynthesizerTrn.infer (models.py:1679-1704).

    scales = [noise_scale, length_scale, noise_scale_w] (onnx_export.py:61-64).
    eps_dp: [B,2,T_x] replaces torch.randn (:96); eps_z: [B,C,>=T_y] or a callable(shape)
    replaces torch.randn_like (:1700)."""
    noise_scale, length_scale, noise_scale_w = [float(s) for s in scales]
    g = None
    if cfg["n_speakers"] > 0:
        g = F.embedding(sid, w["emb_g.weight"]).unsqueeze(-1)
    x, m_p, logs_p, x_mask = text_encoder(tokens, lengths, g, w, cfg)
    logw = sdp_reverse(x, x_mask, g, eps_dp, noise_scale_w, w, cfg)
    w_ceil, y_lengths = durations(logw, x_mask, length_scale)
    idx = frame_to_token(w_ceil, y_lengths)
    Ty = idx.shape[1]
    y_mask = sequence_mask(y_lengths, Ty).unsqueeze(1).to(x_mask.dtype)
    attn = path_from_index(idx, x.shape[2])
    m_e = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)       # :1696
    logs_e = torch.matmul(attn.squeeze(1), logs_p.transpose(1, 2)).transpose(1, 2)
    if callable(eps_z):
        e = eps_z(tuple(m_e.shape))
    else:
        e = eps_z[:, :, :Ty]
    z_p = m_e + e * torch.exp(logs_e) * noise_scale
    z = flow_reverse(z_p, y_mask, g, w, cfg)
    if not decode:      # (long utterances: the caller vocodes slices of z itself, the decoder being local -- +-24 frames)
        return dict(w_ceil=w_ceil, y_lengths=y_lengths, idx=idx, logw=logw, z_p=z_p, z=z, y_mask=y_mask, g=g)
    zin = z * y_mask
    if cfg["decoder"] in ("mb_istft", "ms_istft", "istft"):
        o, o_mb = decoder_mb_istft(zin, w, cfg)
    else:
        o, o_mb = decoder_hifigan(zin, w, cfg, g)
    res = dict(o=o, o_mb=o_mb, w_ceil=w_ceil, y_lengths=y_lengths, idx=idx)
    if return_all:
        res.update(x=x, m_p=m_p, logs_p=logs_p, logw=logw, z_p=z_p, z=z, attn=attn, y_mask=y_mask, g=g)
    return res


def infer_single(w, cfg, tokens, sid, scales, eps_dp, eps_z):
    """B=1 convenience wrapper (numpy in / numpy out) used by tests and the CPU baseline."""
    t = torch.as_tensor(tokens, dtype=torch.long).reshape(1, -1)
    with torch.no_grad():
        r = infer(w, cfg, t, torch.tensor([t.shape[1]]), torch.tensor([int(sid)]), scales,
                  torch.as_tensor(eps_dp).reshape(1, 2, -1), eps_z)
    return r
