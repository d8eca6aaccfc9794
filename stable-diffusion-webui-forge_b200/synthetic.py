"""Synthetic (random-init) checkpoints with the reference's parameter names and shapes, generated directly on
the device — there is no network for real checkpoints, and the benchmark contract asks for random weights of
the named architecture.  Variance-preserving scales keep activations O(1) through the network so that fp16
arithmetic is exercised in its normal range.
"""
from __future__ import annotations

from typing import Dict

import torch

from .unet_engine import unet_structure

# LDM-format configs (huggingface_guess @ 84826248, restated in SURVEY.md §8c)
SD15 = dict(
    in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2], channel_mult=[1, 2, 4, 4],
    transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0], transformer_depth_output=[1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0],
    transformer_depth_middle=1, num_heads=8, num_head_channels=-1, use_spatial_transformer=True,
    use_linear_in_transformer=False, context_dim=768, adm_in_channels=None, num_classes=None)
SDXL = dict(
    in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2], channel_mult=[1, 2, 4],
    transformer_depth=[0, 0, 2, 2, 10, 10], transformer_depth_output=[0, 0, 0, 2, 2, 2, 10, 10, 10],
    transformer_depth_middle=10, num_heads=-1, num_head_channels=64, use_spatial_transformer=True,
    use_linear_in_transformer=True, context_dim=2048, adm_in_channels=2816, num_classes="sequential")
VAE_SDXL = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                latent_channels=4, scaling_factor=0.13025, shift_factor=0.0)

VAE_SD15 = dict(VAE_SDXL, scaling_factor=0.18215)
# Flux / SD3 VAE (backend/huggingface/black-forest-labs/FLUX.1-dev/vae/config.json): 16 latent channels, shift, no post-quant conv
VAE_FLUX = dict(VAE_SDXL, latent_channels=16, scaling_factor=0.3611, shift_factor=0.1159, use_post_quant_conv=False)

# algorithmic FLOPs (2*MACs of conv/linear/attention matmuls) per sample-forward, measured on the reference
# modules with torch.utils.flop_counter
UNET_GFLOP_PER_SAMPLE = {"sdxl@128": 6761.2, "sd15@64": 803.3}
VAE_GFLOP_PER_IMAGE = {"sdxl@1024": 10470.4}
FLUX_GFLOP_PER_SAMPLE = 69466.6  # Flux.1-dev forward at 4096 img + 256 txt tokens


class _Gen:
    def __init__(self, device, dtype, seed):
        self.g = torch.Generator(device=device).manual_seed(seed)
        self.device, self.dtype = device, dtype
        self.sd: Dict[str, torch.Tensor] = {}

    def _randn(self, *shape, scale=1.0):
        return (torch.randn(*shape, generator=self.g, device=self.device, dtype=torch.float32) * scale).to(self.dtype)

    def lin(self, p, cin, cout, bias=True):
        self.sd[p + ".weight"] = self._randn(cout, cin, scale=cin ** -0.5)
        if bias:
            self.sd[p + ".bias"] = self._randn(cout, scale=0.05)

    def conv(self, p, cin, cout, k):
        self.sd[p + ".weight"] = self._randn(cout, cin, k, k, scale=(cin * k * k) ** -0.5)
        self.sd[p + ".bias"] = self._randn(cout, scale=0.05)

    def norm(self, p, c):
        self.sd[p + ".weight"] = (1.0 + self._randn(c, scale=0.1).float()).to(self.dtype)
        self.sd[p + ".bias"] = self._randn(c, scale=0.05)


def random_unet_state_dict(cfg: dict, device="cuda", dtype=torch.float16, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = _Gen(torch.device(device), dtype, seed)
    mc = cfg["model_channels"]
    ted = mc * 4
    g.lin("time_embed.0", mc, ted)
    g.lin("time_embed.2", ted, ted)
    if cfg.get("num_classes") == "sequential":
        g.lin("label_emb.0.0", cfg["adm_in_channels"], ted)
        g.lin("label_emb.0.2", ted, ted)
    st = unet_structure(cfg)
    ctx = cfg["context_dim"]
    use_lin = cfg["use_linear_in_transformer"]
    for blk in st["input"] + [st["middle"]] + st["output"]:
        for layer in blk:
            kind, p = layer[0], layer[1]
            if kind == "conv":
                g.conv(p, layer[2], layer[3], 3)
            elif kind == "res":
                cin, cout = layer[2], layer[3]
                g.norm(p + ".in_layers.0", cin)
                g.conv(p + ".in_layers.2", cin, cout, 3)
                g.lin(p + ".emb_layers.1", ted, cout)
                g.norm(p + ".out_layers.0", cout)
                g.conv(p + ".out_layers.3", cout, cout, 3)
                if cin != cout:
                    g.conv(p + ".skip_connection", cin, cout, 1)
            elif kind == "attn":
                ch, depth = layer[2], layer[5]
                g.norm(p + ".norm", ch)
                if use_lin:
                    g.lin(p + ".proj_in", ch, ch)
                    g.lin(p + ".proj_out", ch, ch)
                else:
                    g.conv(p + ".proj_in", ch, ch, 1)
                    g.conv(p + ".proj_out", ch, ch, 1)
                for d in range(depth):
                    q = f"{p}.transformer_blocks.{d}"
                    for a, kv in (("attn1", ch), ("attn2", ctx)):
                        g.lin(f"{q}.{a}.to_q", ch, ch, bias=False)
                        g.lin(f"{q}.{a}.to_k", kv, ch, bias=False)
                        g.lin(f"{q}.{a}.to_v", kv, ch, bias=False)
                        g.lin(f"{q}.{a}.to_out.0", ch, ch)
                    for n in ("norm1", "norm2", "norm3"):
                        g.norm(f"{q}.{n}", ch)
                    g.lin(f"{q}.ff.net.0.proj", ch, ch * 8)
                    g.lin(f"{q}.ff.net.2", ch * 4, ch)
            elif kind == "down":
                g.conv(p + ".op", layer[2], layer[2], 3)
            elif kind == "up":
                g.conv(p + ".conv", layer[2], layer[2], 3)
    g.norm("out.0", st["out_ch"])
    g.conv("out.2", mc, cfg["out_channels"], 3)
    return g.sd


def random_vae_decoder_state_dict(cfg: dict, device="cuda", dtype=torch.bfloat16, seed: int = 1):
    g = _Gen(torch.device(device), dtype, seed)
    boc = list(cfg["block_out_channels"])
    ch = boc[0]
    ch_mult = [c // ch for c in boc]
    nres, nrb, zc = len(boc), cfg["layers_per_block"], cfg["latent_channels"]

    def res(p, cin, cout):
        g.norm(p + ".norm1", cin)
        g.conv(p + ".conv1", cin, cout, 3)
        g.norm(p + ".norm2", cout)
        g.conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            g.conv(p + ".nin_shortcut", cin, cout, 1)

    if cfg.get("use_post_quant_conv", True):
        g.conv("post_quant_conv", zc, zc, 1)
    block_in = ch * ch_mult[-1]
    g.conv("decoder.conv_in", zc, block_in, 3)
    res("decoder.mid.block_1", block_in, block_in)
    g.norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        g.conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = ch * ch_mult[lvl]
        for j in range(nrb + 1):
            res(f"decoder.up.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            g.conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    g.norm("decoder.norm_out", block_in)
    g.conv("decoder.conv_out", block_in, cfg["out_channels"], 3)
    return g.sd


FLUX_DEV = dict(in_channels=16, vec_in_dim=768, context_in_dim=4096, hidden_size=3072, mlp_ratio=4.0, num_heads=24,
                depth=19, depth_single_blocks=38, axes_dim=[16, 56, 56], theta=10000, qkv_bias=True, guidance_embed=True)


def random_flux_state_dict(cfg: dict, device="cuda", dtype=torch.bfloat16, seed: int = 2) -> Dict[str, torch.Tensor]:
    """Random Flux transformer weights with the reference's parameter names (backend/nn/flux.py:326-353), generated on
    the device (Flux.1-dev: 11.9 B parameters = 23.8 GB in bf16).  Modulation weights are scaled down so (1 + scale)
    stays near 1 and the gates near 0.1-0.3: the residual stream keeps O(1) magnitude through 57 blocks."""
    g = _Gen(torch.device(device), dtype, seed)
    hs, H = cfg["hidden_size"], cfg["num_heads"]
    D = hs // H
    mlp = int(hs * cfg["mlp_ratio"])

    def lin(p, cin, cout, bias=True, wscale=1.0, bmean=0.0):
        g.sd[p + ".weight"] = g._randn(cout, cin, scale=wscale * cin ** -0.5)
        if bias:
            g.sd[p + ".bias"] = (bmean + g._randn(cout, scale=0.05).float()).to(dtype)

    def qknorm(p):
        g.sd[p + ".query_norm.scale"] = (1.0 + g._randn(D, scale=0.1).float()).to(dtype)
        g.sd[p + ".key_norm.scale"] = (1.0 + g._randn(D, scale=0.1).float()).to(dtype)

    lin("img_in", cfg["in_channels"] * 4, hs)
    for name, cin in [("time_in", 256), ("vector_in", cfg["vec_in_dim"])] + ([("guidance_in", 256)] if cfg["guidance_embed"] else []):
        lin(name + ".in_layer", cin, hs)
        lin(name + ".out_layer", hs, hs)
    lin("txt_in", cfg["context_in_dim"], hs)
    for i in range(cfg["depth"]):
        p = f"double_blocks.{i}"
        for s in ("img", "txt"):
            lin(f"{p}.{s}_mod.lin", hs, 6 * hs, wscale=0.3, bmean=0.1)
            lin(f"{p}.{s}_attn.qkv", hs, 3 * hs, bias=cfg["qkv_bias"])
            qknorm(f"{p}.{s}_attn.norm")
            lin(f"{p}.{s}_attn.proj", hs, hs)
            lin(f"{p}.{s}_mlp.0", hs, mlp)
            lin(f"{p}.{s}_mlp.2", mlp, hs)
    for i in range(cfg["depth_single_blocks"]):
        p = f"single_blocks.{i}"
        lin(p + ".linear1", hs, 3 * hs + mlp)
        lin(p + ".linear2", hs + mlp, hs)
        qknorm(p + ".norm")
        lin(p + ".modulation.lin", hs, 3 * hs, wscale=0.3, bmean=0.1)
    lin("final_layer.linear", hs, 4 * cfg["in_channels"])
    lin("final_layer.adaLN_modulation.1", hs, 2 * hs, wscale=0.3)
    return g.sd
