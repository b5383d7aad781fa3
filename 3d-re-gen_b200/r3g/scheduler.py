"""FlowMatchEulerDiscreteScheduler, interface of Hunyuan3D-2/hy3dgen/shapegen/schedulers.py:55-318
(reversed-time Euler: sigmas run 0 -> 1).  Pure host logic; the update itself is the fused
r3g_cfg_euler_step kernel when the pipeline drives it, `step()` is kept for API parity."""
import types

import numpy as np
import torch


class FlowMatchEulerDiscreteSchedulerOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False):
        if use_dynamic_shifting:
            raise NotImplementedError("use_dynamic_shifting is not used by Hunyuan3D-2 shape checkpoints")
        self.config = types.SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift,
                                            use_dynamic_shifting=False)
        t = torch.from_numpy(np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32).copy())
        sig = t / num_train_timesteps
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = sig * num_train_timesteps
        self.sigmas = sig
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self._step_index = None
        self._begin_index = None
        self.num_inference_steps = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        """schedulers.py:181-221.  Kept on the HOST (the per-step scalars feed kernel arguments)."""
        n = self.config.num_train_timesteps
        if sigmas is None:
            self.num_inference_steps = num_inference_steps
            sigmas = np.linspace(self.sigma_max * n, self.sigma_min * n, num_inference_steps) / n
        sigmas = np.asarray(sigmas)
        sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        s = torch.from_numpy(sigmas).to(dtype=torch.float32)
        self.timesteps = s * n
        self.sigmas = torch.cat([s, torch.ones(1)])
        self.num_inference_steps = len(self.timesteps)
        self._step_index = None

    def index_for_timestep(self, timestep):
        idx = (self.timesteps == timestep).nonzero()
        return idx[1 if len(idx) > 1 else 0].item()

    def step(self, model_output, timestep, sample, return_dict=True, **kwargs):
        """schedulers.py:245-318 on torch tensors (API parity; the pipeline uses the fused kernel instead)."""
        if isinstance(timestep, int):
            raise ValueError("pass one of scheduler.timesteps, not an integer index")
        if self._step_index is None:
            self._step_index = self.index_for_timestep(torch.as_tensor(timestep).cpu())
        s0, s1 = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        prev = (sample.to(torch.float32) + (s1 - s0).to(sample.device) * model_output).to(model_output.dtype)
        self._step_index += 1
        return FlowMatchEulerDiscreteSchedulerOutput(prev) if return_dict else (prev,)

    def __len__(self):
        return self.config.num_train_timesteps
