"""Engine configuration = the ``model`` block of the reference training JSON
(/root/reference/training/vits2/configs/mb_istft_vits2_multi.json:42-79) plus
``n_vocab`` (training/vits2/text/symbols.py: 62 symbols) and ``n_speakers`` (json :37).

Constants that the reference hard-codes in ``SynthesizerTrn.__init__``
(training/vits2/models.py:1613-1625) are spelled out here: flow kernel 5, dilation 1,
4 WN layers, 4 flows; SDP filter 256, kernel 3, 4 flows (3 used in reverse, :94-95).
"""
import copy
import json

DEFAULT_CONFIG = {
    "n_vocab": 62,
    "n_speakers": 200,
    "gin_channels": 256,
    "inter_channels": 192,
    "hidden_channels": 192,
    "filter_channels": 768,
    "n_heads": 2,
    "n_layers": 6,
    "kernel_size": 3,
    "window_size": 4,
    "use_spk_conditioned_encoder": True,
    "cond_layer_idx": 2,
    "use_transformer_flows": True,
    "transformer_flow_type": "pre_conv2",
    "flow_n_heads": 2,               # heads of the flow's pre_transformer: hard-coded in the reference (models.py:355), NOT n_heads
    "flow_kernel_size": 5,
    "flow_dilation_rate": 1,
    "flow_wn_layers": 4,
    "flow_n_flows": 4,
    "dp_filter_channels": 256,
    "dp_kernel_size": 3,
    "dp_n_flows": 4,
    "dp_num_bins": 10,
    "dp_tail_bound": 5.0,
    "decoder": "mb_istft",           # "mb_istft" (Multiband_iSTFT_Generator) | "ms_istft" (Multistream_iSTFT_Generator) |
                                     # "istft" (iSTFT_Generator) | "hifigan" (Generator)
    "resblock": "1",
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "upsample_rates": [4, 4],
    "upsample_initial_channel": 512,
    "upsample_kernel_sizes": [16, 16],
    "subbands": 4,
    "gen_istft_n_fft": 16,
    "gen_istft_hop_size": 4,
    "sampling_rate": 22050,
}


ISTFT_DECODERS = ("mb_istft", "ms_istft", "istft")      # decoders that end in conv_post -> exp / pi*sin -> inverse STFT


def from_training_json(path_or_dict, n_vocab=62):
    """Map a reference training config (json :42-79, :37) onto the engine config."""
    cfg = path_or_dict
    if not isinstance(cfg, dict):
        with open(path_or_dict) as f:
            cfg = json.load(f)
    m = cfg["model"]
    out = copy.deepcopy(DEFAULT_CONFIG)
    out["n_vocab"] = n_vocab
    out["n_speakers"] = cfg["data"].get("n_speakers", 0)
    out["sampling_rate"] = cfg["data"].get("sampling_rate", 22050)
    for k in ("gin_channels", "inter_channels", "hidden_channels", "filter_channels", "n_heads",
              "n_layers", "kernel_size", "resblock", "resblock_kernel_sizes",
              "resblock_dilation_sizes", "upsample_rates", "upsample_initial_channel",
              "upsample_kernel_sizes", "subbands", "gen_istft_n_fft", "gen_istft_hop_size"):
        if k in m:
            out[k] = m[k]
    out["use_spk_conditioned_encoder"] = bool(m.get("use_spk_conditioned_encoder", False))
    out["use_transformer_flows"] = bool(m.get("use_transformer_flows", False))
    # reference defaults (models.py:1561-1564): note the flow type defaults to "mono_layer_post_residual"
    out["transformer_flow_type"] = m.get("transformer_flow_type", "mono_layer_post_residual")
    # Only what the engine implements is accepted, with the reason up front instead of a KeyError inside pack():
    #  * use_sdp=False builds the deterministic DurationPredictor (models.py:1625-1628): not implemented;
    #  * use_transformer_flows=True needs "pre_conv2" (ResidualCouplingTransformersLayer2, models.py:329-396);
    #  * use_transformer_flows=False is the plain ResidualCouplingLayer + Flip stack ONLY when the flow type is not
    #    "mono_layer_post_residual": with that (default!) type the reference appends a MonoTransformerFlowLayer to
    #    every flow (models.py:716-734), which the engine does not have.
    if not bool(m.get("use_sdp", True)):
        raise ValueError("use_sdp=false (deterministic DurationPredictor, models.py:1627) is not supported by this engine")
    if out["use_transformer_flows"]:
        if out["transformer_flow_type"] != "pre_conv2":
            raise ValueError("transformer_flow_type %r not supported (only 'pre_conv2')" % out["transformer_flow_type"])
    elif out["transformer_flow_type"] == "mono_layer_post_residual":
        raise ValueError("use_transformer_flows=false with transformer_flow_type 'mono_layer_post_residual' (the reference "
                         "default) adds MonoTransformerFlowLayers to the flow (models.py:716-734): not supported")
    # same precedence as SynthesizerTrn.__init__ (models.py:1585-1606)
    if m.get("mb_istft_vits", False):
        out["decoder"] = "mb_istft"
    elif m.get("ms_istft_vits", False):
        out["decoder"] = "ms_istft"
    elif m.get("istft_vits", False):
        out["decoder"] = "istft"
        out["subbands"] = 1                     # iSTFT_Generator has a single band and no synthesis filter bank
    else:
        out["decoder"] = "hifigan"
    return out


def hop_total(cfg):
    """Output samples per latent frame (256 for the reference config)."""
    up = 1
    for u in cfg["upsample_rates"]:
        up *= u
    if cfg["decoder"] in ISTFT_DECODERS:
        up *= cfg["gen_istft_hop_size"] * cfg["subbands"]
    return up
