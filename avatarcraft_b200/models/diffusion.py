"""Score-distillation guidance with the reference's interface (models/diffusion.py:24-339 `StableDiffusion`):
`get_text_embeds`, `mannual_backward`, `calc_grad`, `train_step`, `encode_imgs`, `decode_latents`,
`produce_latents`.  SURVEY.md 8a row S1.

What the reference delegates to third-party packages that are NOT in /root/reference and not installed here
(diffusers==0.16.1, transformers==4.29.1, HF weights -- readme.md:35, models/diffusion.py:53-63) is restated from
scratch: UNet2DConditionModel (models/sd_unet.py), AutoencoderKL (models/sd_vae.py), the scaled-linear DDPM noise
schedule the PNDMScheduler exposes as `alphas_cumprod` (:63-64) and its `add_noise`.  Without weights the networks
are randomly initialised (`weights_dir=None`); with a diffusers-format directory they load by name.  **Parity
unpinned**: there is no third-party oracle in this image; tests pin the wrapper arithmetic (schedule, noise mixing,
classifier-free guidance, w(t), clamp, manual backward) against closed forms and the native kernels against the same
network evaluated with torch ops.

Per step (`mannual_backward`, :92-149): bilinear resize to 512x512 -> VAE encode WITH grad (posterior.sample() *
0.18215) -> t ~ U{20..980} -> noisy latents -> UNet on the (uncond, text) pair, NO grad -- the native tcgen05 path of
sd_ops -> eps_uncond + s (eps_text - eps_uncond) -> grad = clamp((1 - abar_t)(eps - noise), -1, 1) ->
latents.backward(grad)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sd_unet import UNet2DConditionModel, UNetConfig
from .sd_vae import AutoencoderKL


def scaled_linear_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """PNDMScheduler(beta_schedule="scaled_linear").alphas_cumprod (models/diffusion.py:63-64)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class NoiseSchedule:
    """The two things the reference uses of its scheduler in the SDS step: `alphas_cumprod` and `add_noise`."""

    def __init__(self, num_train_timesteps=1000):
        self.alphas_cumprod = scaled_linear_alphas_cumprod(num_train_timesteps)

    def add_noise(self, original, noise, timesteps):
        ac = self.alphas_cumprod.to(original.device)[timesteps].to(original.dtype)
        a = ac.sqrt().reshape(-1, *([1] * (original.dim() - 1)))
        b = (1.0 - ac).sqrt().reshape(-1, *([1] * (original.dim() - 1)))
        return a * original + b * noise


class HashTextEncoder(nn.Module):
    """Stand-in text encoder when no CLIP weights / vocabulary are available offline: a deterministic byte-pair hash
    tokenizer (BOS, tokens, EOS padding to 77, CLIP's framing) feeding an embedding + 2-layer transformer of the right
    output shape [B, 77, dim].  Real CLIP modules can be passed to StableDiffusion(tokenizer=..., text_encoder=...)."""

    def __init__(self, dim=768, vocab=49408, max_length=77):
        super().__init__()
        self.vocab, self.model_max_length = vocab, max_length
        self.token_embedding = nn.Embedding(vocab, dim)
        self.position_embedding = nn.Embedding(max_length, dim)
        layer = nn.TransformerEncoderLayer(dim, 8, dim * 4, dropout=0.0, batch_first=True, norm_first=True)
        self.encoder = nn.TransformerEncoder(layer, 2)
        self.final_layer_norm = nn.LayerNorm(dim)

    def tokenize(self, prompts):
        ids = torch.full((len(prompts), self.model_max_length), self.vocab - 1, dtype=torch.long)     # EOS = pad (CLIP)
        for i, p in enumerate(prompts):
            toks = [self.vocab - 2]                                                                   # BOS
            for w in p.lower().split():
                h = 2166136261
                for ch in w.encode():
                    h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
                toks.append(h % (self.vocab - 2))
            toks = toks[:self.model_max_length - 1] + [self.vocab - 1]
            ids[i, :len(toks)] = torch.tensor(toks)
        return ids

    def forward(self, ids):
        L = ids.shape[1]
        x = self.token_embedding(ids) + self.position_embedding(torch.arange(L, device=ids.device))[None]
        mask = torch.full((L, L), float("-inf"), device=ids.device).triu(1)
        return (self.final_layer_norm(self.encoder(x, mask=mask)),)


class _NativeVaeMoments(torch.autograd.Function):
    """quant_conv(encoder(x)) of the frozen VAE with both directions on the native kernels (models/sd_vae_native.py): the
    autograd graph of the SDS step (models/diffusion.py:104-148) only sees this one node between the bilinear resize and the
    posterior sample.  The incoming gradient is scaled by a power of two (device scalar, no sync) so that its fp16 GEMM
    operands keep their precision; the backward is linear in it."""

    @staticmethod
    def forward(ctx, x, vae):
        from . import sd_vae_native
        moments, ctx.backward_fn = sd_vae_native.encode_moments(vae, x.detach())
        return moments

    @staticmethod
    def backward(ctx, g):
        gmax = g.abs().amax().clamp_min(1e-30)
        scale = torch.exp2(torch.floor(torch.log2(64.0 / gmax)).clamp(-60.0, 60.0))
        return ctx.backward_fn(g * scale) / scale, None


class StableDiffusion(nn.Module):
    def __init__(self, device, version="1.5", weights_dir=None, unet=None, vae=None, text_encoder=None, tokenizer=None,
                 unet_config=None, seed=0):
        super().__init__()
        self.sd_version, self.device = version, torch.device(device)
        self.num_train_timesteps = 1000
        self.min_step = int(self.num_train_timesteps * 0.02)
        self.max_step = int(self.num_train_timesteps * 0.98)
        self.use_depth = version == "2.0"
        if version not in ("1.5", "2.0", "2.1"):
            raise ValueError("sd_version must be '1.5' or '2.0' (models/diffusion.py:45-49), or '2.1' (BASELINE config 5)")
        cfg = unet_config or (UNetConfig.sd2_depth() if self.use_depth else UNetConfig.sd21() if version == "2.1" else UNetConfig.sd15())
        on_cuda = self.device.type == "cuda"
        # modules are created (and randomly initialised) directly on the target device: 0.94 G parameters take seconds
        # on the GPU and most of a minute on the host; the seed makes the random networks identical on every rank
        with torch.random.fork_rng(devices=[self.device] if on_cuda else []), torch.device(self.device):
            torch.manual_seed(seed)
            self.vae = vae if vae is not None else AutoencoderKL()
            self.unet = unet if unet is not None else UNet2DConditionModel(cfg)
            self.text_encoder = text_encoder if text_encoder is not None else HashTextEncoder(cfg.cross_attention_dim)
        self.tokenizer = tokenizer
        # multi-GPU: split the classifier-free-guidance pair over ranks.  Only valid when every rank holds bit-identical
        # (latents, t, noise): the trainer guarantees that through pixel_gradient(seed=...) on the all-gathered pass-1 image
        # and switches it on; anywhere else the ranks draw their own randoms, so it is off by default.
        self.cfg_parallel = False
        self.native_vae = True          # VAE encoder forward + backward on the native kernels (CUDA only); False = torch autograd ops
        self._user_text_encoder = text_encoder is not None
        if weights_dir is not None:
            self._load_diffusers_dir(weights_dir)
        self.to(self.device)
        for p in self.parameters():
            p.requires_grad_(False)                          # guidance networks are frozen; gradients flow to the image only
        self.scheduler = NoiseSchedule(self.num_train_timesteps)
        self.alphas = self.scheduler.alphas_cumprod.to(self.device)

    def _load_diffusers_dir(self, root):
        """diffusers layout: <root>/{unet,vae}/diffusion_pytorch_model.{safetensors,bin} and the text side
        <root>/text_encoder + <root>/tokenizer (models/diffusion.py:53-60 loads all four from one model id).  The prompt
        conditions the UNet through the text encoder: real UNet weights with the random stand-in encoder would make
        SDS silently ignore the prompt, so a missing text side is an error unless the caller passed its own modules."""
        def load_state(sub):
            st = os.path.join(root, sub, "diffusion_pytorch_model.safetensors")
            if os.path.exists(st):
                from safetensors.torch import load_file
                return load_file(st, device="cpu")
            f = os.path.join(root, sub, "diffusion_pytorch_model.bin")
            if not os.path.exists(f):
                raise FileNotFoundError(f"{st} (or .bin)")
            return torch.load(f, map_location="cpu")
        self.unet.load_state_dict(load_state("unet"))
        self.vae.load_state_dict(load_state("vae"))
        if self._user_text_encoder and self.tokenizer is not None:
            return
        te, tk = os.path.join(root, "text_encoder"), os.path.join(root, "tokenizer")
        if not (os.path.isdir(te) and os.path.isdir(tk)):
            raise FileNotFoundError(f"{te} and {tk}: the checkpoint's CLIP text encoder and tokenizer are required with real "
                                    "UNet weights (or pass text_encoder= and tokenizer=)")
        from transformers import CLIPTextModel, CLIPTokenizer
        self.tokenizer = CLIPTokenizer.from_pretrained(tk)
        self.text_encoder = CLIPTextModel.from_pretrained(te)

    # ---- text ----------------------------------------------------------------------------------------------------
    def get_text_embeds(self, prompt):
        """[uncond ; text] embeddings [2B, 77, D] (models/diffusion.py:72-89)."""
        if not isinstance(prompt, list):
            prompt = [prompt]

        def embed(texts):
            if self.tokenizer is not None:
                ids = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                                     return_tensors="pt").input_ids
            else:
                ids = self.text_encoder.tokenize(texts)
            with torch.no_grad():
                return self.text_encoder(ids.to(self.device))[0]
        return torch.cat([embed([""] * len(prompt)), embed(prompt)])

    # ---- the SDS step ----------------------------------------------------------------------------------------------
    def encode_imgs(self, imgs):
        """imgs [B,3,H,W] in [0,1] -> latents [B,4,H/8,W/8] = posterior.sample() * 0.18215 (:304-312)."""
        if imgs.is_cuda and self.native_vae:          # both directions of the encoder on the sm_100a kernels
            from .sd_vae import DiagonalGaussian
            posterior = DiagonalGaussian(_NativeVaeMoments.apply(2 * imgs - 1, self.vae))
        else:                                          # torch ops (CPU unit tests, A/B)
            posterior = self.vae.encode(2 * imgs - 1).latent_dist
        return posterior.sample() * 0.18215

    def decode_latents(self, latents):
        with torch.no_grad():
            imgs = self.vae.decode(latents / 0.18215).sample
        return (imgs / 2 + 0.5).clamp(0, 1)

    def sds_latent_gradient(self, latents, text_embeddings, t, noise, guidance_scale=100, pred_depth=None):
        """grad = clamp(w(t) (eps_cfg - noise), -1, 1) with w = 1 - abar_t (:121-146); no autograd."""
        with torch.no_grad():
            latents_noisy = self.scheduler.add_noise(latents, noise, t)
            import torch.distributed as dist
            if self.cfg_parallel and dist.is_available() and dist.is_initialized() and dist.get_world_size() >= 2:
                # the (uncond, text) pair is the one piece of the SD step that shards: even ranks evaluate
                # the unconditional half, odd ranks the text half, one all-gather of the [1,4,64,64] prediction (64 KB)
                half = dist.get_rank() % 2
                x = latents_noisy
                if self.use_depth and pred_depth is not None:
                    x = torch.cat([x, pred_depth[:1]], dim=1)
                mine = self.unet(x, t, encoder_hidden_states=text_embeddings[half:half + 1]).sample.contiguous()
                parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
                dist.all_gather(parts, mine)
                noise_pred_uncond, noise_pred_text = parts[0], parts[1]
            else:
                latent_model_input = torch.cat([latents_noisy] * 2)
                if self.use_depth and pred_depth is not None:
                    latent_model_input = torch.cat([latent_model_input, pred_depth], dim=1)
                noise_pred = self.unet(latent_model_input, t, encoder_hidden_states=text_embeddings).sample
                noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
            noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
            w = 1 - self.alphas[t]
            return (w * (noise_pred - noise)).clamp(-1, 1)

    def mannual_backward(self, text_embeddings, pred_rgb: torch.Tensor, guidance_scale=100, pred_depth: torch.Tensor = None) -> None:
        """Back-propagates the SDS gradient into `pred_rgb` ([1,3,H,W], requires grad) -- models/diffusion.py:92-149.
        (The reference's zero-pad result is immediately overwritten by the bilinear resize, :104-105: only the resize
        has an effect.)"""
        pred_rgb_512 = F.interpolate(pred_rgb, (512, 512), mode="bilinear", align_corners=False)
        if self.use_depth and pred_depth is not None:
            pred_depth = F.interpolate(pred_depth, size=(64, 64), mode="bicubic", align_corners=False)
            pred_depth = 2.0 * (pred_depth - pred_depth.min()) / (pred_depth.max() - pred_depth.min()) - 1.0
            pred_depth = torch.cat([pred_depth] * 2)
        t = torch.randint(self.min_step, self.max_step + 1, [1], dtype=torch.long, device=self.device)
        latents = self.encode_imgs(pred_rgb_512)                       # requires grad
        noise = torch.randn_like(latents)
        grad = self.sds_latent_gradient(latents.detach(), text_embeddings, t, noise, guidance_scale, pred_depth)
        latents.backward(gradient=grad, retain_graph=True)

    def pixel_gradient(self, text_embeddings, rgb_rows, h, w, guidance_scale=100, seed=None):
        """Trainer-facing form of the step (stylize.py:120-131): rgb_rows [h*w,3] (the no-grad pass-1 render) ->
        d(SDS)/d(rgb) [h*w,3].  `seed` makes the step's random draws (t, posterior sample, noise) identical on every
        rank of a multi-GPU job, so replicated VAE work agrees bit for bit and no broadcast is needed."""
        if seed is not None:
            torch.manual_seed(int(seed))
        elif self.cfg_parallel and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            raise RuntimeError("cfg_parallel needs identical (t, noise) on every rank: pass seed= (or set cfg_parallel = False)")
        img = rgb_rows.detach().reshape(1, h, w, 3).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        with torch.enable_grad():
            self.mannual_backward(text_embeddings, img, guidance_scale)
        return img.grad[0].permute(1, 2, 0).reshape(h * w, 3).contiguous()

    def calc_grad(self, text_embeddings, pred_rgb: torch.Tensor, guidance_scale=100) -> torch.Tensor:
        """Same step, returning d(SDS)/d(pred_rgb) as a tensor (:154-216)."""
        self.mannual_backward(text_embeddings, pred_rgb, guidance_scale)
        return pred_rgb.grad.detach().clone()

    def train_step(self, text_embeddings, pred_rgb, guidance_scale=100):
        self.mannual_backward(text_embeddings, pred_rgb, guidance_scale)
        return 0                                                        # dummy loss value, as the reference (:264)

    def produce_latents(self, text_embeddings, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, latents=None):
        """Deterministic DDIM sampling over the same schedule (the reference drives diffusers' PNDM stepper, :266-289;
        the multistep PNDM update is scheduler code outside this path's scope)."""
        if latents is None:
            latents = torch.randn((text_embeddings.shape[0] // 2, self.unet.in_channels, height // 8, width // 8), device=self.device)
        ts = torch.linspace(self.num_train_timesteps - 1, 0, num_inference_steps, device=self.device).long()
        with torch.no_grad():
            for i, t in enumerate(ts):
                eps = self.unet(torch.cat([latents] * 2), t, encoder_hidden_states=text_embeddings)["sample"]
                eu, et = eps.chunk(2)
                eps = eu + guidance_scale * (et - eu)
                a_t = self.alphas[t]
                a_prev = self.alphas[ts[i + 1]] if i + 1 < len(ts) else torch.ones((), device=self.device)
                x0 = (latents - (1 - a_t).sqrt() * eps) / a_t.sqrt()
                latents = a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * eps
        return latents

    def prompt_to_img(self, prompts, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, latents=None):
        if isinstance(prompts, str):
            prompts = [prompts]
        text_embeds = self.get_text_embeds(prompts)
        latents = self.produce_latents(text_embeds, height=height, width=width, latents=latents,
                                       num_inference_steps=num_inference_steps, guidance_scale=guidance_scale)
        imgs = self.decode_latents(latents).detach().cpu().permute(0, 2, 3, 1).numpy()
        return (imgs * 255).round().astype("uint8")
