"""DDPM ancestral sampler with the reference's API (sd/ddpm.py:5-186).

The schedule tables are host-side fp32 tensors exactly as in the reference (including its
beta_start = 0.000085 default, sd/ddpm.py:30). `step()` on CUDA tensors runs the fused
sdb_cfg_ddpm_step kernel; `coefficient_table()` exposes the per-step constants the CUDA-graph
loop in pipeline.generate() reads from device memory.
"""
import numpy as np
import torch


class DDPMSampler:
    def __init__(self, generator: torch.Generator, num_training_steps=1000,
                 beta_start: float = 0.000085, beta_end: float = 0.012):
        # scaled-linear schedule in fp32 (sd/ddpm.py:43-48)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_training_steps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.generator = generator
        self.num_train_timesteps = num_training_steps
        self.timesteps = torch.from_numpy(np.arange(0, num_training_steps)[::-1].copy())

    # ------------------------------------------------------------------ schedule (sd/ddpm.py:56-99)
    def set_inference_timesteps(self, num_inference_steps=50):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps)

    def _get_previous_timestep(self, timestep: int) -> int:
        return timestep - self.num_train_timesteps // self.num_inference_steps

    def _get_variance(self, timestep: int) -> torch.Tensor:
        prev_t = self._get_previous_timestep(timestep)
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        current_beta_t = 1 - alpha_prod_t / alpha_prod_t_prev
        variance = (1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t
        return torch.clamp(variance, min=1e-20)

    def set_strength(self, strength=1):
        start_step = self.num_inference_steps - int(self.num_inference_steps * strength)
        self.timesteps = self.timesteps[start_step:]
        self.start_step = start_step

    # ------------------------------------------------------------------ per-step constants
    def step_coefficients(self, timestep):
        """(sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt, sigma_t) as fp32 0-dim tensors — the
        scalars sd/ddpm.py:107-133 computes, in the same fp32 operation order."""
        t = int(timestep)
        prev_t = self._get_previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        c_x0 = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
        c_xt = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        sigma = self._get_variance(t) ** 0.5 if t > 0 else torch.tensor(0.0)
        return beta_prod_t ** 0.5, alpha_prod_t ** 0.5, c_x0, c_xt, sigma

    def coefficient_table(self, device=None):
        """fp32 [len(timesteps), 5] table of step_coefficients for the current timestep list."""
        rows = [torch.stack([c.to(torch.float32) for c in self.step_coefficients(t)])
                for t in self.timesteps]
        table = torch.stack(rows).contiguous()
        return table.to(device) if device is not None else table

    # ------------------------------------------------------------------ sd/ddpm.py:102-139
    def step(self, timestep: int, latents: torch.Tensor, model_output: torch.Tensor):
        t = int(timestep)
        sb, sa, c_x0, c_xt, sigma = self.step_coefficients(t)
        noise = None
        if t > 0:
            noise = torch.randn(model_output.shape, generator=self.generator,
                                device=model_output.device, dtype=model_output.dtype)
        if latents.is_cuda:
            from . import ops
            coef = torch.stack([sb, sa, c_x0, c_xt, sigma]).to(torch.float32).view(1, 5).to(latents.device)
            out = latents.to(torch.float32).contiguous().clone()
            ops.cfg_ddpm_step(out, model_output.to(torch.float32).contiguous(), noise, coef, 0, 1.0,
                              False, None, eps_nchw=True)
            return out
        pred_original_sample = (latents - sb * model_output) / sa
        pred_prev_sample = c_x0 * pred_original_sample + c_xt * latents
        if noise is not None:
            pred_prev_sample = pred_prev_sample + sigma * noise
        return pred_prev_sample

    # ------------------------------------------------------------------ sd/ddpm.py:143-186
    def add_noise(self, original_samples: torch.FloatTensor, timesteps: torch.IntTensor) -> torch.FloatTensor:
        alphas_cumprod = self.alphas_cumprod.to(dtype=original_samples.dtype)
        t = timesteps.to("cpu") if torch.is_tensor(timesteps) else torch.tensor(timesteps)
        sqrt_alpha_prod = (alphas_cumprod[t] ** 0.5).flatten()
        sqrt_one_minus_alpha_prod = ((1 - alphas_cumprod[t]) ** 0.5).flatten()
        noise = torch.randn(original_samples.shape, generator=self.generator,
                            device=original_samples.device, dtype=original_samples.dtype)
        if original_samples.is_cuda and sqrt_alpha_prod.numel() == 1:
            from . import ops
            return ops.axpby(original_samples.contiguous(), noise, float(sqrt_alpha_prod[0]),
                             float(sqrt_one_minus_alpha_prod[0]))
        sa = sqrt_alpha_prod.to(original_samples.device)
        sb = sqrt_one_minus_alpha_prod.to(original_samples.device)
        while len(sa.shape) < len(original_samples.shape):
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise
