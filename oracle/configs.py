"""ORACLE — test infrastructure only.

LDM-format UNet / VAE configurations of the reference.  The reference obtains them from the third-party
package `huggingface_guess` @ 84826248b49bb7ca754c73293299c4d4e23a548d (modules/launch_utils.py:397,404;
imported at backend/loader.py:7,452), which is not vendored under /root/reference; the values below are
restated from SURVEY.md §8c, which checked them against the vendored diffusers configs
(backend/huggingface/*/unet/config.json via backend/misc/diffusers_state_dict.py:70-134) and the canonical
parameter counts (SD1.5 859.5 M, SDXL 2567.5 M).

`TINY_*` are reduced-width configurations with the same block structure, used for CPU-sized parity tests
and golden vectors; every channel count stays a multiple of 64 so the same kernels are exercised.
"""

SD15 = dict(
    in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2], channel_mult=[1, 2, 4, 4],
    transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0], transformer_depth_output=[1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0],
    transformer_depth_middle=1, num_heads=8, num_head_channels=-1, use_spatial_transformer=True,
    use_linear_in_transformer=False, context_dim=768, adm_in_channels=None, num_classes=None,
)

SDXL = dict(
    in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2], channel_mult=[1, 2, 4],
    transformer_depth=[0, 0, 2, 2, 10, 10], transformer_depth_output=[0, 0, 0, 2, 2, 2, 10, 10, 10],
    transformer_depth_middle=10, num_heads=-1, num_head_channels=64, use_spatial_transformer=True,
    use_linear_in_transformer=True, context_dim=2048, adm_in_channels=2816, num_classes="sequential",
)

# same topology as SDXL (3 levels, linear proj, label_emb, head dim 64), 64/128/256 channels
TINY_XL = dict(
    in_channels=4, out_channels=4, model_channels=64, num_res_blocks=[2, 2, 2], channel_mult=[1, 2, 4],
    transformer_depth=[0, 0, 1, 1, 2, 2], transformer_depth_output=[0, 0, 0, 1, 1, 1, 2, 2, 2],
    transformer_depth_middle=2, num_heads=-1, num_head_channels=64, use_spatial_transformer=True,
    use_linear_in_transformer=True, context_dim=128, adm_in_channels=96, num_classes="sequential",
)

# same topology as SD1.5 (4 levels, conv proj, no label_emb) but head dim 64 (the fused path's head dims)
TINY_15 = dict(
    in_channels=4, out_channels=4, model_channels=64, num_res_blocks=[1, 1, 1, 1], channel_mult=[1, 2, 4, 4],
    transformer_depth=[1, 1, 1, 0], transformer_depth_output=[1, 1, 1, 1, 1, 1, 0, 0],
    transformer_depth_middle=1, num_heads=-1, num_head_channels=64, use_spatial_transformer=True,
    use_linear_in_transformer=False, context_dim=128, adm_in_channels=None, num_classes=None,
)

# SD1.5-style head split (num_heads fixed -> head dims 8/16/32 here, 40/80/160 at full width): exercises the
# head-dim padding of the fused path
TINY_15H = dict(TINY_15, num_heads=8, num_head_channels=-1)

# SD2.x topology (4 levels, linear proj, head dim 64, no label_emb; v-prediction for the 768 models):
# huggingface_guess SD20 unet_config restated: context_dim 1024, num_head_channels 64, use_linear_in_transformer True
SD21 = dict(SD15, num_heads=-1, num_head_channels=64, use_linear_in_transformer=True, context_dim=1024)
TINY_21 = dict(TINY_15, use_linear_in_transformer=True)

# SDXL refiner (backend/huggingface/stabilityai/stable-diffusion-xl-refiner-1.0/unet/config.json: block_out_channels
# 384/768/1536/1536, cross-attention in levels 1 and 2 with 4 transformer layers, head dim 64, linear projections,
# cross_attention_dim 1280, projection_class_embeddings_input_dim 2560) in the LDM form the reference's loader builds
SDXL_REFINER = dict(
    in_channels=4, out_channels=4, model_channels=384, num_res_blocks=[2, 2, 2, 2], channel_mult=[1, 2, 4, 4],
    transformer_depth=[0, 0, 4, 4, 4, 4, 0, 0], transformer_depth_output=[0, 0, 0, 4, 4, 4, 4, 4, 4, 0, 0, 0],
    transformer_depth_middle=4, num_heads=-1, num_head_channels=64, use_spatial_transformer=True,
    use_linear_in_transformer=True, context_dim=1280, adm_in_channels=2560, num_classes="sequential",
)

# SDXL VAE (backend/huggingface/stabilityai/stable-diffusion-xl-base-1.0/vae/config.json)
VAE_SDXL = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                latent_channels=4, scaling_factor=0.13025, shift_factor=0.0)
VAE_SD15 = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                latent_channels=4, scaling_factor=0.18215, shift_factor=0.0)
TINY_VAE = dict(in_channels=3, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
                latent_channels=4, scaling_factor=0.13025, shift_factor=0.0)

# Flux / SD3 VAE (backend/huggingface/black-forest-labs/FLUX.1-dev/vae/config.json): 16 latent channels, a shift factor,
# no quant / post-quant convolutions
VAE_FLUX = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                latent_channels=16, scaling_factor=0.3611, shift_factor=0.1159, use_post_quant_conv=False)
TINY_VAE_FLUX = dict(TINY_VAE, latent_channels=16, scaling_factor=0.3611, shift_factor=0.1159, use_post_quant_conv=False)

CONFIGS = {"sd15": SD15, "sdxl": SDXL, "sd21": SD21, "sdxl_refiner": SDXL_REFINER, "tiny_xl": TINY_XL, "tiny_15": TINY_15, "tiny_15h": TINY_15H, "tiny_21": TINY_21}
VAE_CONFIGS = {"sdxl": VAE_SDXL, "sd15": VAE_SD15, "tiny": TINY_VAE, "flux": VAE_FLUX, "tiny_flux": TINY_VAE_FLUX}
