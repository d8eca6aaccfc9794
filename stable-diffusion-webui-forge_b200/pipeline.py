"""Public API of the backend: the txt2img denoise job (what `StableDiffusionProcessingTxt2Img.sample` +
`decode_latent_batch` do in the reference, modules/processing.py:1342-1428, 628-635) driven entirely through
the sm_100a kernels.

    pipe = Txt2ImgPipeline(unet_cfg, unet_state_dict, vae_cfg=..., vae_state_dict=...)
    out = pipe.generate(cond, uncond, noise=..., steps=30, sampler="euler_a", cfg_scale=7.0)

Inputs are the same objects the reference's sampler receives: conditioning dicts {"crossattn": [B,77,ctx],
"vector": [B,adm]} (backend/sampling/condition.py:91-119), initial noise N(0,1) [B,4,h,w] (ImageRNG,
modules/rng.py) and — for ancestral samplers — one noise tensor per step.  Host tensors are accepted
(pinned or not) and copied in; the UNet forward is captured once per (batch, latent size) in a CUDA graph
and replayed every step.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import ops, sampling
from .unet_engine import UNetEngine


class GraphedUNet:
    """One CUDA graph of `UNetEngine.forward_sigma` for fixed shapes, with static input/output buffers."""

    def __init__(self, engine: UNetEngine, batch: int, reps: int, hh: int, ww: int, n_ctx: int, use_graph: bool = True):
        self.engine = engine
        dev, dt = engine.device, engine.dtype
        cfg = engine.cfg
        n = batch * reps
        self.x = torch.zeros((batch, cfg["in_channels"], hh, ww), dtype=torch.float32, device=dev)
        self.sigma = torch.ones((batch,), dtype=torch.float32, device=dev)
        self.timesteps = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.context = torch.zeros((n, n_ctx, cfg["context_dim"]), dtype=dt, device=dev)
        self.y = torch.zeros((n, cfg["adm_in_channels"]), dtype=dt, device=dev) if engine.has_label else None
        self.reps = reps
        # cross-attention K|V projections of the (per-job constant) context: filled once per job, read by the graph
        self.kv_cache = engine.alloc_kv_cache(n, n_ctx)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.eps: Optional[torch.Tensor] = None
        self.launches_per_forward = 0
        if use_graph:
            self._capture()

    def _eager(self) -> torch.Tensor:
        return self.engine.forward_sigma(self.x, self.sigma, self.timesteps, self.context, self.y, self.reps,
                                         kv_cache=self.kv_cache)

    def set_context(self, context: torch.Tensor, y: Optional[torch.Tensor]) -> None:
        """New conditioning for a job: copy in, then project K|V of every cross-attention layer once."""
        self.context.copy_(context)
        if self.y is not None:
            self.y.copy_(y)
        self.engine.fill_kv_cache(self.context, self.kv_cache)

    def _capture(self) -> None:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: first-call attribute setup, allocator pools
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = ops.LAUNCHES
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.eps = self._eager()
        self.launches_per_forward = ops.LAUNCHES - before
        self.graph = g

    def __call__(self) -> torch.Tensor:
        if self.graph is None:
            before = ops.LAUNCHES
            self.eps = self._eager()
            self.launches_per_forward = ops.LAUNCHES - before
        else:
            self.graph.replay()
            ops.LAUNCHES += self.launches_per_forward
        return self.eps


class Txt2ImgPipeline:
    def __init__(self, unet_cfg: dict, unet_state_dict: Dict[str, torch.Tensor], *, vae_cfg: Optional[dict] = None,
                 vae_state_dict: Optional[Dict[str, torch.Tensor]] = None, dtype: torch.dtype = torch.float16,
                 vae_dtype: torch.dtype = torch.bfloat16, device="cuda", use_graph: bool = True,
                 vae_encoder_state_dict: Optional[Dict[str, torch.Tensor]] = None, prediction_type: str = "epsilon"):
        self.device = torch.device(device)
        self.dtype = dtype
        self.unet = UNetEngine(unet_cfg, unet_state_dict, dtype=dtype, device=device)
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"prediction_type {prediction_type!r}: epsilon (SD1.x, SDXL) or v_prediction (SD2.x 768-v)")
        self.pred = sampling.Prediction(prediction_type=prediction_type)  # k_prediction.py:113-167
        self.use_graph = use_graph
        self._graphs: Dict[tuple, GraphedUNet] = {}
        self.vae = None
        if vae_cfg is not None:
            from .vae_engine import VAEDecoderEngine
            self.vae = VAEDecoderEngine(vae_cfg, vae_state_dict, dtype=vae_dtype, device=device)
        self.vae_encoder = None
        if vae_cfg is not None and vae_encoder_state_dict is not None:
            from .vae_engine import VAEEncoderEngine
            self.vae_encoder = VAEEncoderEngine(vae_cfg, vae_encoder_state_dict, dtype=vae_dtype, device=device)

    def _graph_for(self, batch, reps, hh, ww, n_ctx) -> GraphedUNet:
        key = (batch, reps, hh, ww, n_ctx)
        if not self.unet.supports_latent(hh, ww):
            raise ops.B200Error(-2, f"latent size {hh}x{ww} does not tile on the TMA convolution path "
                                    "(every UNet level needs a width that is a power of two <= 128 or a multiple of 128)")
        if key not in self._graphs:
            self._graphs[key] = GraphedUNet(self.unet, batch, reps, hh, ww, n_ctx, self.use_graph)
        return self._graphs[key]

    @torch.no_grad()
    def sample(self, cond: dict, uncond: Optional[dict], noise: torch.Tensor, *, steps: int = 30,
               sampler: str = "euler_a", cfg_scale: float = 7.0, sigmas: Optional[torch.Tensor] = None,
               step_noise: Optional[torch.Tensor] = None, eta: float = 1.0, s_noise: float = 1.0,
               callback: Optional[Callable] = None, init_latent: Optional[torch.Tensor] = None,
               sgm_noise_multiplier: bool = False) -> torch.Tensor:
        """Returns the final latent [B,4,h,w] fp32 on the device.
        sgm_noise_multiplier: Forge's option of that name (modules/shared_options.py:410, default False) — the txt2img start
        is noise * sigmas[0] by default and noise * sqrt(1 + sigmas[0]^2) with it (max_denoise, k_prediction.py:94-104).
        init_latent (img2img): the start is noise * sigmas[0] + init_latent (noise_scaling with max_denoise=False,
        k_prediction.py:94-104) instead of the txt2img start.
        noise: [B,4,h,w] N(0,1) (host or device).  step_noise: [steps-1 or more, B,4,h,w] for ancestral samplers
        (row i is added after step i, matching the order in which ImageRNG.next() is drawn by the reference)."""
        dev = self.device
        b, c, hh, ww = noise.shape
        has_uncond = uncond is not None and not abs(cfg_scale - 1.0) < 1e-9  # sampling_function.py:295-298
        reps = 2 if has_uncond else 1
        n_ctx = cond["crossattn"].shape[1]
        gu = self._graph_for(b, reps, hh, ww, n_ctx)
        if sigmas is None:
            sigmas = sampling.make_sigmas(self.pred, sampler, steps)
        sigmas = sigmas.float().cpu()
        builder = sampling.SAMPLERS[sampler][0]
        plan = builder(sigmas, eta, s_noise) if sampler == "euler_a" else builder(sigmas)

        # ---- host -> device: conditioning, initial noise, per-step noise
        def put(t, dtype):
            return t.to(device=dev, dtype=dtype, non_blocking=True)

        ctx = [cond["crossattn"]] if not has_uncond else [uncond["crossattn"], cond["crossattn"]]  # [uncond, cond]
        yv = None
        if gu.y is not None:
            ys = [cond["vector"]] if not has_uncond else [uncond["vector"], cond["vector"]]
            yv = torch.cat([put(t, self.dtype) for t in ys], 0)
        gu.set_context(torch.cat([put(t, self.dtype) for t in ctx], 0), yv)
        x = gu.x
        # modules/sd_samplers_kdiffusion.py:207 -> k_prediction.py:94-104 (txt2img: zero latent, max_denoise = the option)
        if init_latent is not None:
            init_latent = put(init_latent, torch.float32)
            if init_latent.data_ptr() == x.data_ptr():  # a latent returned by a previous sample() IS this buffer
                init_latent = init_latent.clone()
        x.copy_(put(noise, torch.float32))
        if init_latent is None:
            x.mul_(float(torch.sqrt(1.0 + sigmas[0] ** 2.0)) if sgm_noise_multiplier else float(sigmas[0]))
        else:
            x.mul_(float(sigmas[0])).add_(init_latent)
        sn_dev = put(step_noise, torch.float32) if step_noise is not None else None
        # per-step scalar tables (sigma per image, UNet timestep = index of nearest log-sigma)
        sig_tab = sigmas[:-1].to(dev).view(-1, 1).expand(-1, b).contiguous()
        t_tab = self.pred.timestep(sigmas[:-1]).float().to(dev).view(-1, 1).expand(-1, b * reps).contiguous()

        def eps_fn(i):
            gu.sigma.copy_(sig_tab[i])
            gu.timesteps.copy_(t_tab[i])
            return gu()

        noise_fn = (lambda i: sn_dev[i]) if sn_dev is not None else None
        if sampling.SAMPLERS[sampler][2] and noise_fn is None:
            raise ValueError(f"sampler {sampler} needs step_noise")
        sampling.run_sampler(eps_fn, x, plan, cfg_scale=cfg_scale, has_uncond=has_uncond, noise_fn=noise_fn,
                             callback=callback, prediction=1 if self.pred.prediction_type == "v_prediction" else 0)
        return x.clone()  # x is the graph's static input buffer: the caller gets its own copy

    @staticmethod
    def img2img_schedule(sigmas: torch.Tensor, steps: int, denoising_strength: float) -> torch.Tensor:
        """setup_img2img_steps (modules/sd_samplers_common.py:24-33, default options) + the slice of
        sample_img2img (modules/sd_samplers_kdiffusion.py:140-143): the last t_enc + 1 sigmas of the full schedule."""
        t_enc = int(min(denoising_strength, 0.999) * steps)
        return sigmas[steps - t_enc - 1:]

    @torch.no_grad()
    def img2img(self, cond: dict, uncond: Optional[dict], init: torch.Tensor, noise: torch.Tensor, *, steps: int = 30,
                denoising_strength: float = 0.75, sampler: str = "euler_a", cfg_scale: float = 7.0,
                vae_noise: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
        """img2img / hires-fix second pass (KDiffusionSampler.sample_img2img, modules/sd_samplers_kdiffusion.py:136-194):
        `init` is either pixels NHWC [B,H,W,3] in [0,1] (encoded through the fused VAE encoder + process_in, as
        StableDiffusionProcessingImg2Img.init does via images_tensor_to_samples) or a latent NCHW [B,4,h,w].
        Returns the final latent."""
        if init.dim() == 4 and init.shape[-1] == 3:
            if self.vae_encoder is None:
                raise RuntimeError("pipeline built without a VAE encoder")
            latent = self.vae_encoder.encode(init.to(self.device).float().contiguous(), vae_noise, process_in=True)
        else:
            latent = init
        full = sampling.make_sigmas(self.pred, sampler, steps)
        sched = self.img2img_schedule(full, steps, denoising_strength)
        return self.sample(cond, uncond, noise, steps=len(sched) - 1, sampler=sampler, cfg_scale=cfg_scale, sigmas=sched,
                           init_latent=latent, **kw)

    @torch.no_grad()
    def hires_fix(self, cond: dict, uncond: Optional[dict], noise: torch.Tensor, noise_hr: torch.Tensor, *, steps: int = 30,
                  hr_steps: Optional[int] = None, denoising_strength: float = 0.7, latent_mode: str = "bilinear",
                  antialias: bool = False, sampler: str = "euler_a", cfg_scale: float = 7.0,
                  step_noise: Optional[torch.Tensor] = None, step_noise_hr: Optional[torch.Tensor] = None) -> torch.Tensor:
        """txt2img with a latent hires-fix second pass (StableDiffusionProcessingTxt2Img.sample + sample_hr_pass,
        modules/processing.py:1342-1428, 1430-1540, latent upscalers only): first pass at noise.shape, the latent is
        resized to noise_hr.shape with torch.nn.functional.interpolate exactly as the reference does (:1458 — Forge calls
        torch there too, it is not part of the hot path), then img2img over the last t_enc + 1 sigmas of an
        `hr_steps or steps` schedule.  Both passes run the same fused denoise path."""
        first = self.sample(cond, uncond, noise, steps=steps, sampler=sampler, cfg_scale=cfg_scale, step_noise=step_noise)
        up = torch.nn.functional.interpolate(first, size=tuple(noise_hr.shape[2:]), mode=latent_mode,
                                             **({} if latent_mode == "nearest" else {"antialias": antialias}))
        return self.img2img(cond, uncond, up, noise_hr, steps=hr_steps or steps, denoising_strength=denoising_strength,
                            sampler=sampler, cfg_scale=cfg_scale, step_noise=step_noise_hr)

    @torch.no_grad()
    def decode(self, latent: torch.Tensor) -> torch.Tensor:
        """latent [B,4,h,w] fp32 -> images [B,H,W,3] fp32 in [0,1] (VAE.decode, backend/patcher/vae.py:128-155)."""
        if self.vae is None:
            raise RuntimeError("pipeline built without a VAE")
        return self.vae.decode(latent)

    @torch.no_grad()
    def generate(self, cond: dict, uncond: Optional[dict], noise: torch.Tensor, **kw) -> torch.Tensor:
        latent = self.sample(cond, uncond, noise, **kw)
        return self.decode(latent) if self.vae is not None else latent


class GraphedFlux:
    """One CUDA graph of `FluxEngine.forward_nhwc` for fixed shapes, with static input/output buffers."""

    def __init__(self, engine, batch: int, hh: int, ww: int, n_txt: int, use_graph: bool = True):
        self.engine = engine
        dev, dt, cfg = engine.device, engine.dtype, engine.cfg
        self.x = torch.zeros((batch, cfg["in_channels"], hh, ww), dtype=torch.float32, device=dev)
        self.t = torch.ones((batch,), dtype=torch.float32, device=dev)
        self.guidance = torch.full((batch,), 3.5, dtype=torch.float32, device=dev)
        self.context = torch.zeros((batch, n_txt, cfg["context_in_dim"]), dtype=dt, device=dev)
        self.y = torch.zeros((batch, cfg["vec_in_dim"]), dtype=dt, device=dev)
        self.out = torch.zeros((batch, hh, ww, cfg["in_channels"]), dtype=dt, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_forward = 0
        if use_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            before = ops.LAUNCHES
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager()
            self.launches_per_forward = ops.LAUNCHES - before
            self.graph = g

    def _eager(self) -> torch.Tensor:
        return self.engine.forward_nhwc(self.x, self.t, self.context, self.y,
                                        self.guidance if self.engine.guidance_embed else None, out=self.out)

    def __call__(self) -> torch.Tensor:
        if self.graph is None:
            before = ops.LAUNCHES
            self._eager()
            self.launches_per_forward = ops.LAUNCHES - before
        else:
            self.graph.replay()
            ops.LAUNCHES += self.launches_per_forward
        return self.out


class FluxTxt2ImgPipeline:
    """Flux txt2img denoise job: what the reference runs through KModel.apply_model with PredictionFlux
    (backend/modules/k_model.py:25-46, k_prediction.py:285-322), CFG scale 1 (distilled guidance enters the model as an
    embedding, backend/diffusion_engine/flux.py), Euler over the "Simple" schedule.  One fused launch per sampler step."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], *, dtype: torch.dtype = torch.bfloat16, device="cuda",
                 use_graph: bool = True, vae_cfg: Optional[dict] = None, vae_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 vae_dtype: torch.dtype = torch.bfloat16):
        from .flux_engine import FluxEngine
        self.device, self.dtype = torch.device(device), dtype
        self.model = FluxEngine(cfg, state_dict, dtype=dtype, device=device)
        self.use_graph = use_graph
        self._graphs: Dict[tuple, GraphedFlux] = {}
        self.vae = None
        if vae_cfg is not None:  # the 16-channel Flux VAE (process_out = z / 0.3611 + 0.1159) on the same fused decoder
            from .vae_engine import VAEDecoderEngine
            self.vae = VAEDecoderEngine(vae_cfg, vae_state_dict, dtype=vae_dtype, device=device)

    @torch.no_grad()
    def decode(self, latent: torch.Tensor) -> torch.Tensor:
        if self.vae is None:
            raise RuntimeError("pipeline built without a VAE")
        return self.vae.decode(latent)

    @torch.no_grad()
    def generate(self, cond: dict, noise: torch.Tensor, **kw) -> torch.Tensor:
        """sample + VAE decode: images [B, H, W, 3] fp32 in [0, 1] (the latent when no VAE was given)."""
        latent = self.sample(cond, noise, **kw)
        return self.decode(latent) if self.vae is not None else latent

    def _graph_for(self, batch, hh, ww, n_txt) -> GraphedFlux:
        key = (batch, hh, ww, n_txt)
        if key not in self._graphs:
            self._graphs[key] = GraphedFlux(self.model, batch, hh, ww, n_txt, self.use_graph)
        return self._graphs[key]

    @torch.no_grad()
    def sample(self, cond: dict, noise: torch.Tensor, *, steps: int = 20, guidance: float = 3.5,
               sigmas: Optional[torch.Tensor] = None, callback: Optional[Callable] = None,
               init_latent: Optional[torch.Tensor] = None) -> torch.Tensor:
        """cond = {"crossattn": [B, Lt, ctx] (T5), "vector": [B, vec] (pooled CLIP)}; noise [B, 16, h, w] N(0,1).
        init_latent (img2img): the start is sigma_0 * noise + (1 - sigma_0) * init_latent ('const' noise_scaling,
        k_prediction.py:94-96).  Returns the final latent [B, 16, h, w] fp32 on the device."""
        dev = self.device
        b, c, hh, ww = noise.shape
        gf = self._graph_for(b, hh, ww, cond["crossattn"].shape[1])
        if sigmas is None:
            pred = sampling.FluxPrediction(seq_len=(hh // 2) * (ww // 2))
            sigmas = sampling.get_sigmas_simple(pred.sigmas, steps)
        sigmas = sigmas.float().cpu()
        plan = sampling.plan_euler(sigmas)
        gf.context.copy_(cond["crossattn"].to(device=dev, dtype=self.dtype, non_blocking=True))
        gf.y.copy_(cond["vector"].to(device=dev, dtype=self.dtype, non_blocking=True))
        gf.guidance.fill_(float(guidance))
        x = gf.x
        if init_latent is not None:
            init_latent = init_latent.to(device=dev, dtype=torch.float32).clone()
        x.copy_(noise.to(device=dev, dtype=torch.float32, non_blocking=True))
        x.mul_(float(sigmas[0]))  # 'const' noise_scaling (k_prediction.py:94-96): sigma * noise + (1 - sigma) * latent
        if init_latent is not None:
            x.add_(init_latent.mul_(1.0 - float(sigmas[0])))
        sig_tab = sigmas[:-1].to(dev).view(-1, 1).expand(-1, b).contiguous()

        def model_fn(i):
            gf.t.copy_(sig_tab[i])  # PredictionFlux.timestep(sigma) = sigma (k_prediction.py:311-312)
            return gf()

        sampling.run_sampler(model_fn, x, plan, cfg_scale=1.0, has_uncond=False, callback=callback)
        return x.clone()  # x is the graph's static input buffer: the caller gets its own copy


    @torch.no_grad()
    def img2img(self, cond: dict, init_latent: torch.Tensor, noise: torch.Tensor, *, steps: int = 20,
                denoising_strength: float = 0.75, guidance: float = 3.5, **kw) -> torch.Tensor:
        """img2img on a Flux latent: the last t_enc + 1 sigmas of the Simple schedule (setup_img2img_steps,
        modules/sd_samplers_common.py:24-33; sample_img2img, modules/sd_samplers_kdiffusion.py:140-146)."""
        b, c, hh, ww = noise.shape
        pred = sampling.FluxPrediction(seq_len=(hh // 2) * (ww // 2))
        full = sampling.get_sigmas_simple(pred.sigmas, steps)
        sched = Txt2ImgPipeline.img2img_schedule(full, steps, denoising_strength)
        return self.sample(cond, noise, steps=len(sched) - 1, guidance=guidance, sigmas=sched, init_latent=init_latent, **kw)
