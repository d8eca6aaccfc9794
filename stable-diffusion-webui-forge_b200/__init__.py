"""b200forge — a B200-native (sm_100a) diffusion inference backend that plugs in under
stable-diffusion-webui-forge's backend/ surface (attention_function, ForgeOperations,
model_function_wrapper, k-diffusion samplers, VAE decode wrapper).

The directory is named after the reference repository (`stable-diffusion-webui-forge_b200`), which is not
a valid Python identifier; import it as `b200forge` (a two-line alias package at the repo root).
"""
__version__ = "0.1.0"
