"""Factorised-prior ``EntropyBottleneck`` on the B200 kernels.

Mirror of the compressai==1.1.5 classes the reference imports
(``compressai.entropy_models.EntropyBottleneck`` / ``EntropyModel`` and
``compressai.models.utils.update_registered_buffers``; reference call sites
hub/compressor.py:12-13,49-63,98-100,124 and lossyless/rates.py:10-18,55-71): same
constructor, parameter / buffer names (so ``hub/*/factorized_rate.pt`` load unchanged),
``forward / compress / decompress / update / loss`` semantics and error behaviour.

What runs where:
  * ``forward`` (likelihood) -> one fused CUDA launch (csrc/eb_likelihood.cu) with a custom
    autograd backward; ``compress`` / ``decompress`` -> batched rANS kernels (csrc/codec.cu).
  * ``update`` builds the integer CDF tables at load time: the pmf with float32 torch CPU ops
    in compressai's op order (the reference also does this on the CPU, hub/compressor.py:63),
    the pmf -> CDF quantisation in the library's host function ``ll_build_cdf_tables``.
  * ``loss`` (aux loss over the 3 quantiles per channel, 1536 elements) stays plain torch.
There is no CPU path for forward / compress / decompress: tensors must be on a B200.
"""

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N

__all__ = ["EntropyBottleneck", "GaussianConditional", "EntropyModel", "LowerBound", "update_registered_buffers"]


class LowerBound(nn.Module):
    """compressai.ops.LowerBound: max(x, bound) whose gradient passes where x >= bound or the
    gradient pushes x up.  Kept as a module so the ``likelihood_lower_bound.bound`` buffer of
    reference checkpoints loads; the clamp itself is fused into the likelihood kernel."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)).type_as(g) * g, None


class _RansCoder:
    """Stands in for compressai's ``_EntropyCoder("ans")`` handle (dropped / restored by
    ``make_pickable_`` / ``undo_pickable_``, lossyless/rates.py:273-284).  Stateless: the codec
    lives in liblossyless_b200.so."""

    name = "ans"

    def __init__(self):
        N.load()  # fail loudly if the extension is missing


def default_entropy_coder():
    return "ans"


class EntropyModel(nn.Module):
    """compressai.entropy_models.EntropyModel surface used by the reference."""

    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        if entropy_coder not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{entropy_coder}" (available: ans)')
        if int(entropy_coder_precision) != 16:
            raise ValueError("only entropy_coder_precision=16 is supported")
        self.entropy_coder = None  # created lazily: constructing a model must work without a GPU
        self.entropy_coder_precision = 16
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._enc_cache = None  # (key, device tensors) derived from the CDF buffers

    # -- compressai API -------------------------------------------------------------
    def quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    @staticmethod
    def dequantize(inputs, means=None):
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.float()
        return outputs

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """[C, Lmax] pmf + [C, 1] tail -> int32 [C, Lmax + 2] (host function of the library)."""
        pmf = pmf.detach().float().cpu().contiguous()
        tail = tail_mass.detach().float().cpu().reshape(-1).contiguous()
        plen = pmf_length.detach().int().cpu().contiguous()
        cdf = torch.zeros((pmf.shape[0], max_length + 2), dtype=torch.int32)
        N.call("ll_build_cdf_tables", N.ptr(pmf), N.ptr(tail), N.ptr(plen), pmf.shape[0], int(max_length), N.ptr(cdf))
        return cdf


    # -- entropy coding -----------------------------------------------------------------
    def _coder_tables(self, device):
        """Device-resident tables for the rANS kernels, cached until the CDF buffers change."""
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, str(device))
        if self._enc_cache is not None and self._enc_cache[0] == key:
            return self._enc_cache[1]
        cdf = self._quantized_cdf.to(device=device, dtype=torch.int32).contiguous()
        cdf_len = self._cdf_length.to(device=device, dtype=torch.int32).contiguous()
        offset = self._offset.to(device=device, dtype=torch.int32).contiguous()
        C, stride = cdf.shape
        enc = torch.empty(C * stride * 16, dtype=torch.uint8, device=device)
        N.call("ll_rans_build_enc_table", N.ptr(cdf), N.ptr(cdf_len), C, stride, N.ptr(enc), N.stream())
        tabs = dict(cdf=cdf, cdf_len=cdf_len, offset=offset, enc=enc, stride=stride)
        self._enc_cache = (key, tabs)
        return tabs

    def _channel_means(self, device):
        """Per-channel mean of the prior (the bottleneck's medians); conditional priors pass `means`."""
        raise ValueError("this entropy model has no per-channel means: pass `means`")

    def compress_to_device(self, z, bias=None, exp_scale=None, mode=0, want_symbols=False, indexes=None, means=None):
        """Fused (process_z_in +) quantise + rANS encode + pack of z [rows, C] on the GPU.
        Returns (packed uint8 [total], offsets int64 [rows+1], lens int32 [rows], symbols|None),
        all device tensors.  mode 0 = raw concatenation, 1 = container records (">I len" + bytes).
        ``indexes`` int32 [rows, C] / ``means`` float32 [rows, C] select a CDF row and a mean per element
        (conditional priors); by default element c uses CDF row c and the prior's per-channel mean."""
        N.require_cuda(z, "z")
        if self.entropy_coder is None:
            self.entropy_coder = _RansCoder()
        z = z.contiguous()
        rows, C = z.shape
        if indexes is None and C != getattr(self, "channels", C):
            raise ValueError(f"expected {self.channels} channels, got {C}")
        t = self._coder_tables(z.device)
        med = self._channel_means(z.device) if means is None else None
        idx = indexes.to(torch.int32).contiguous() if indexes is not None else None
        mu = means.detach().float().contiguous() if means is not None else None
        if rows == 0:
            return (torch.empty(0, dtype=torch.uint8, device=z.device), torch.zeros(1, dtype=torch.int64, device=z.device),
                    torch.empty(0, dtype=torch.int32, device=z.device),
                    torch.empty((0, C), dtype=torch.int32, device=z.device) if want_symbols else None)
        cap = N.load().ll_rans_stream_cap_bytes(C)
        scratch = torch.empty((rows, cap), dtype=torch.uint8, device=z.device)
        lens = torch.empty(rows, dtype=torch.int32, device=z.device)
        symbols = torch.empty((rows, C), dtype=torch.int32, device=z.device) if want_symbols else None
        b = bias.detach().float().contiguous() if bias is not None else None
        es = exp_scale.detach().float().contiguous() if exp_scale is not None else None
        N.call("ll_rans_encode_indexed", N.ptr(z), N.dtype_code(z), rows, C, N.ptr(b), N.ptr(es), N.ptr(med), N.ptr(idx), N.ptr(mu),
               N.ptr(t["enc"]), t["stride"], N.ptr(t["cdf_len"]), N.ptr(t["offset"]), N.ptr(scratch), N.ptr(lens),
               N.ptr(symbols), N.stream())
        packed = torch.empty(rows * (cap + 4) + 16, dtype=torch.uint8, device=z.device)
        offsets = torch.empty(rows + 1, dtype=torch.int64, device=z.device)
        N.call("ll_rans_pack", N.ptr(scratch), N.ptr(lens), rows, cap, mode, N.ptr(packed), N.ptr(offsets), N.stream())
        return packed, offsets, lens, symbols

    def decompress_from_device(self, data, offsets, lens, bias=None, exp_scale=None, want_symbols=False, indexes=None, means=None,
                               status_out=None, out=None):
        """Batched decode of streams stored in a device byte buffer -> (z_hat [rows,C] fp32,
        symbols|None).  Raises ValueError if any stream is malformed -- unless ``status_out`` (int32 [rows], device)
        is given: then the per-row status words are left there for the caller to check once (no host
        synchronisation per call; this is what lets a decode loop be pipelined).  ``out`` optionally supplies the
        [rows, C] float32 destination."""
        N.require_cuda(data, "data")
        if self.entropy_coder is None:
            self.entropy_coder = _RansCoder()
        t = self._coder_tables(data.device)
        rows = lens.numel()
        C = indexes.shape[1] if indexes is not None else self.channels
        med = self._channel_means(data.device) if means is None else None
        idx = indexes.to(torch.int32).contiguous() if indexes is not None else None
        mu = means.detach().float().contiguous() if means is not None else None
        if rows == 0:
            return torch.empty((0, C), dtype=torch.float32, device=data.device), None
        z_hat = out if out is not None else torch.empty((rows, C), dtype=torch.float32, device=data.device)
        if z_hat.shape != (rows, C) or z_hat.dtype != torch.float32 or not z_hat.is_contiguous():
            raise ValueError("`out` must be a contiguous float32 [rows, C] tensor")
        status = status_out if status_out is not None else torch.empty(rows, dtype=torch.int32, device=data.device)
        symbols = torch.empty((rows, C), dtype=torch.int32, device=data.device) if want_symbols else None
        offsets, lens = offsets.contiguous(), lens.contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        es = exp_scale.detach().float().contiguous() if exp_scale is not None else None
        N.call("ll_rans_decode_indexed", N.ptr(data), N.ptr(offsets), N.ptr(lens), rows, C, N.ptr(t["cdf"]), t["stride"],
               N.ptr(t["cdf_len"]), N.ptr(t["offset"]), N.ptr(med), N.ptr(idx), N.ptr(mu), N.ptr(b), N.ptr(es), N.ptr(z_hat),
               N.ptr(symbols), N.ptr(status), N.stream())
        if status_out is None:
            bad = int(status.sum().item())
            if bad:
                raise ValueError(f"{bad} of {rows} rANS streams are truncated or corrupt")
        return z_hat, symbols

    @staticmethod
    def _strings_to_device(strings, device):
        """list of bytes -> (device byte buffer, offsets int64 [n+1], lens int32 [n]); validates lengths."""
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        lens_np = np.fromiter((len(s) for s in strings), dtype=np.int32, count=len(strings))
        if (lens_np % 4 != 0).any() or (lens_np < 8).any():
            raise ValueError("rANS streams must be >= 8 bytes and a multiple of 4 bytes long")
        off_np = np.zeros(len(strings) + 1, dtype=np.int64)
        np.cumsum(lens_np, out=off_np[1:])
        data = torch.frombuffer(bytearray(b"".join(strings)) + bytearray(16), dtype=torch.uint8).to(device)
        return data, torch.from_numpy(off_np).to(device), torch.from_numpy(lens_np).to(device)

    @staticmethod
    def _device_to_strings(packed, offsets):
        off = offsets.cpu().numpy()
        buf = packed[: int(off[-1])].cpu().numpy().tobytes()
        return [buf[off[i]: off[i + 1]] for i in range(len(off) - 1)]


    # -- checkpoint compatibility ---------------------------------------------------------
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # the CDF buffers are dynamically sized: adopt the checkpoint's shapes (policy "resize")
        for name in ("_quantized_cdf", "_offset", "_cdf_length", "scale_table"):
            key = prefix + name
            if key in state_dict and hasattr(self, name):
                cur = getattr(self, name)
                setattr(self, name, torch.empty(state_dict[key].shape, dtype=cur.dtype, device=cur.device))
        self._enc_cache = None
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class EntropyBottleneck(EntropyModel):
    """compressai.entropy_models.EntropyBottleneck (Balle et al. 2018 factorised prior)."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)

        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))

        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _channel_means(self, device):
        return self._get_medians().detach().reshape(-1).float().to(device).contiguous()

    def _logits_cumulative(self, inputs, stop_gradient):
        """Plain torch: used by update() (load time) and loss() (3 points per channel)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(F.softplus(matrix), logits)
            logits = logits + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def update(self, force=False):
        """Build ``_quantized_cdf / _offset / _cdf_length``.  Always evaluated with float32 CPU
        ops so the tables (hence the bitstream) do not depend on the device the module is on."""
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.quantiles.device
        with torch.no_grad():
            q = self.quantiles.detach().float().cpu()
            medians = q[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max().item())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
            cpu = _CpuParams(self)
            lower = cpu.logits_cumulative(samples - 0.5)
            upper = cpu.logits_cumulative(samples + 0.5)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = (-minima).to(dev)
        self._quantized_cdf = cdf.to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        self._enc_cache = None
        return True

    def loss(self):
        """``sum |logits_cumulative(quantiles) - target|`` with the parameters detached (compressai's auxiliary loss).
        On a B200 with the reference's filters: forward and the gradient w.r.t. the quantiles in one native launch."""
        if self.quantiles.is_cuda and self.filters == (3, 3, 3, 3):
            return _EBAuxLoss.apply(self, self.quantiles)
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    # -- likelihood ---------------------------------------------------------------------
    def forward(self, x, training=None, noise=None):
        """x: [B, C, *spatial] (compressai layout) or [B, C].  Returns (outputs, likelihood).
        ``noise`` (same shape as x, values in (-1/2, 1/2)) overrides the internally drawn
        U(-1/2,1/2) sample in training mode (deterministic tests)."""
        if training is None:
            training = self.training
        shape = x.shape
        if x.dim() < 2 or shape[1] != self.channels:
            raise ValueError(f"expected [B, {self.channels}, ...] input, got {tuple(shape)}")
        if x.dim() > 2:
            if any(s != 1 for s in shape[2:]):
                # B x C x H x W -> rows = B*H*W, channels last
                perm = (0,) + tuple(range(2, x.dim())) + (1,)
                rows = x.permute(*perm).reshape(-1, self.channels)
            else:
                rows = x.reshape(shape[0], self.channels)
        else:
            rows = x
        nz = None
        if training:
            nz = torch.empty_like(rows, dtype=torch.float32).uniform_(-0.5, 0.5) if noise is None else noise.reshape(rows.shape).float()
        out, lik, _ = eb_rate_forward(self, rows, nz, None, None, want_rates=False)
        if x.dim() > 2 and any(s != 1 for s in shape[2:]):
            inv = (0, x.dim() - 1) + tuple(range(1, x.dim() - 1))
            out = out.reshape(shape[0], *shape[2:], self.channels).permute(*inv).contiguous()
            lik = lik.reshape(shape[0], *shape[2:], self.channels).permute(*inv).contiguous()
        else:
            out, lik = out.reshape(shape), lik.reshape(shape)
        return out, lik

    def compress(self, x, bias=None, exp_scale=None):
        """x: [B, C, 1, 1] or [B, C] (already ``process_z_in``-ed unless bias / exp_scale are
        given, in which case that affine map is fused into the kernel) -> list of B bytes."""
        if x.dim() == 4:
            if x.shape[2:] != (1, 1):
                raise ValueError("only 1x1 spatial latents are supported (the reference codes vectors)")
            x = x.reshape(x.shape[0], x.shape[1])
        elif x.dim() != 2:
            raise ValueError("Invalid `inputs` size. Expected a 4-D tensor.")
        packed, offsets, _, _ = self.compress_to_device(x, bias, exp_scale, mode=0)
        return self._device_to_strings(packed, offsets)

    def decompress(self, strings, size=(1, 1), bias=None, exp_scale=None):
        """list of bytes -> [B, C, 1, 1] float32 (compressai layout)."""
        if tuple(size) != (1, 1):
            raise ValueError("only 1x1 spatial latents are supported")
        device = self._quantized_cdf.device
        if device.type != "cuda":
            raise N.NativeError("EntropyBottleneck.decompress needs the module on a B200 (cuda); there is no CPU path")
        data, off, lens = self._strings_to_device(strings, device)
        z_hat, _ = self.decompress_from_device(data, off, lens, bias, exp_scale)
        return z_hat.reshape(len(strings), self.channels, 1, 1)



class GaussianConditional(EntropyModel):
    """compressai.entropy_models.GaussianConditional: Gaussian prior whose per-element scale (and mean)
    is predicted from side information (Balle et al. 2018 / Minnen et al. 2018).  Used by the reference's
    scale-hyperprior rate (lossyless/rates.py:571-756).

    ``update_scale_table`` builds one integer CDF per scale level at load time (float32 CPU ops in
    compressai's op order + the library's host pmf->CDF function); ``compress`` / ``decompress`` are the
    batched rANS kernels with a per-element table index and mean.  The likelihood is a handful of
    elementwise torch ops (erfc), differentiable by autograd -- it is not on the encode path."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs):
        # the complementary error function maximises numerical precision
        return 0.5 * torch.erfc(-(2 ** -0.5) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        import scipy.stats

        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        dev = self.scale_table.device
        with torch.no_grad():
            table = self.scale_table.detach().float().cpu()
            multiplier = -self._standardized_quantile(self.tail_mass / 2)
            pmf_center = torch.ceil(table * multiplier).int()
            pmf_length = 2 * pmf_center + 1
            max_length = int(torch.max(pmf_length).item())
            samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
            samples_scale = table.unsqueeze(1).float()
            upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
            lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
            pmf = upper - lower
            tail_mass = 2 * lower[:, :1]
            cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._quantized_cdf = cdf.to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        self._enc_cache = None

    def _likelihood(self, inputs, scales, means=None):
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None, noise=None):
        """-> (outputs, likelihood).  CUDA float32 inputs take the fused kernel (one launch forward, one backward);
        anything else the elementwise torch ops above (host-side table construction, CPU tests)."""
        out, lik, _ = self.forward_rates(inputs, scales, means, training, noise, want_rates=False)
        return out, lik

    def forward_rates(self, inputs, scales, means=None, training=None, noise=None, want_rates=True):
        """``forward`` plus ``rates = -log(likelihood).sum(-1)`` per row of the last dimension, fused into the same
        launch (what HRateHyperprior.forward_help computes next, lossyless/rates.py:640-646)."""
        if training is None:
            training = self.training
        if inputs.is_cuda and self.use_likelihood_bound:
            nz = None
            if training:
                nz = noise if noise is not None else torch.empty_like(inputs, dtype=torch.float32).uniform_(-0.5, 0.5)
            out, lik, rates = _GCLikelihood.apply(self, inputs, scales, means, nz, bool(want_rates))
            return out, lik, rates
        if training:
            outputs = inputs + (noise if noise is not None else torch.empty_like(inputs).uniform_(-0.5, 0.5))
        else:
            outputs = self.quantize(inputs, "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood, (-torch.log(likelihood).sum(-1) if want_rates else None)

    def build_indexes(self, scales):
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes

    def compress(self, inputs, indexes, means=None, bias=None, exp_scale=None):
        """inputs / indexes / means: [B, C, 1, 1] or [B, C] -> list of B bytes."""
        B = inputs.shape[0]
        x, idx = inputs.reshape(B, -1), indexes.reshape(B, -1)
        mu = means.reshape(B, -1) if means is not None else torch.zeros_like(x, dtype=torch.float32)
        packed, offsets, _, _ = self.compress_to_device(x, bias, exp_scale, mode=0, indexes=idx, means=mu)
        return self._device_to_strings(packed, offsets)

    def decompress(self, strings, indexes, means=None, bias=None, exp_scale=None):
        """-> float32 tensor shaped like ``indexes`` (symbols + means)."""
        device = self._quantized_cdf.device
        if device.type != "cuda":
            raise N.NativeError("GaussianConditional.decompress needs the module on a B200 (cuda); there is no CPU path")
        B = len(strings)
        idx = indexes.reshape(B, -1)
        mu = means.reshape(B, -1) if means is not None else torch.zeros(idx.shape, dtype=torch.float32, device=device)
        data, off, lens = self._strings_to_device(strings, device)
        z_hat, _ = self.decompress_from_device(data, off, lens, bias, exp_scale, indexes=idx, means=mu)
        return z_hat.reshape(indexes.shape)


class _CpuParams:
    """float32 CPU copies of the bottleneck parameters for update()."""

    def __init__(self, eb):
        n = len(eb.filters) + 1
        self.m = [getattr(eb, f"_matrix{i}").detach().float().cpu() for i in range(n)]
        self.b = [getattr(eb, f"_bias{i}").detach().float().cpu() for i in range(n)]
        self.f = [getattr(eb, f"_factor{i}").detach().float().cpu() for i in range(n - 1)]

    def logits_cumulative(self, x):
        logits = x
        for i in range(len(self.m)):
            logits = torch.matmul(F.softplus(self.m[i]), logits)
            logits = logits + self.b[i]
            if i < len(self.f):
                logits = logits + torch.tanh(self.f[i]) * torch.tanh(logits)
        return logits


def update_registered_buffers(module, module_name, buffer_names, state_dict, policy="resize_if_empty", dtype=torch.int):
    """compressai.models.utils.update_registered_buffers: resize ``module``'s (dynamically sized)
    buffers to the shapes found in ``state_dict`` before ``load_state_dict``."""
    valid = [n for n, _ in module.named_buffers()]
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        key = f"{module_name}.{name}"
        if key not in state_dict:
            raise KeyError(key)
        new_size = state_dict[key].size()
        cur = getattr(module, name)
        if policy in ("resize_if_empty", "resize"):
            if policy == "resize" or cur.numel() == 0:
                setattr(module, name, torch.empty(new_size, dtype=dtype, device=cur.device))
        elif policy == "register":
            module.register_buffer(name, torch.empty(new_size, dtype=dtype).fill_(0))
        else:
            raise ValueError(f'Invalid policy "{policy}"')


# ---------------------------------------------------------------------------------------
# fused rate forward / backward
# ---------------------------------------------------------------------------------------
def _packed_params(eb, device):
    """softplus(matrix) / tanh(factor) packed [58, C] on the device (one tiny launch)."""
    if eb.filters != (3, 3, 3, 3):
        raise NotImplementedError("the fused likelihood kernel supports filters=(3,3,3,3) only (the reference's setting)")
    ms = [getattr(eb, f"_matrix{i}").detach().float().contiguous() for i in range(5)]
    bs = [getattr(eb, f"_bias{i}").detach().float().contiguous() for i in range(5)]
    fs = [getattr(eb, f"_factor{i}").detach().float().contiguous() for i in range(4)]
    # [2 * 58, C]: the derived parameters, then their exponent-scaled twins the forward kernel reads (ll_eb_pack_params)
    packed = torch.empty((2 * N.LL_EB_NPARAM, eb.channels), dtype=torch.float32, device=device)
    N.call("ll_eb_pack_params", N.ptr_array(ms), N.ptr_array(bs), N.ptr_array(fs), eb.channels, N.ptr(packed), N.stream())
    return packed, (ms, bs, fs)


def _bound(eb):
    """The LowerBound value as a host float, read back from the (checkpointed) buffer once per buffer version: a
    `.item()` per call is a stream synchronisation in the middle of every training step."""
    if not eb.use_likelihood_bound:
        return 0.0
    buf = eb.likelihood_lower_bound.bound
    key = (buf.data_ptr(), buf._version)
    cache = getattr(eb, "_bound_cache", None)
    if cache is None or cache[0] != key:
        cache = eb._bound_cache = (key, float(buf.item()))
    return cache[1]


class _EBRate(torch.autograd.Function):
    """(z, biasing, scaling, *eb params) -> (z_hat, q, rates) through ll_eb_forward/backward."""

    @staticmethod
    def forward(ctx, eb, z, noise, biasing, scaling, exp_scale, want_rates, *params):
        N.require_cuda(z, "z")
        z = z.contiguous()
        rows, C = z.shape
        dev = z.device
        packed, raw = _packed_params(eb, dev)
        has_affine = biasing is not None
        b = biasing.detach().float().contiguous() if has_affine else None
        es = None
        if has_affine:
            es = exp_scale.detach().float().contiguous() if exp_scale is not None else scaling.detach().float().exp().contiguous()
        med = eb._get_medians().detach().reshape(-1).float().contiguous()
        z_hat = torch.empty((rows, C), dtype=torch.float32, device=dev)
        q = torch.empty((rows, C), dtype=torch.float32, device=dev)
        rates = torch.zeros(rows, dtype=torch.float32, device=dev) if want_rates else None
        nz = noise.contiguous() if noise is not None else None
        N.call("ll_eb_forward", N.ptr(z), N.dtype_code(z), rows, C, N.ptr(nz), N.ptr(b), N.ptr(es), N.ptr(med),
               N.ptr(packed), _bound(eb), N.ptr(z_hat), N.ptr(q), N.ptr(rates), N.stream())
        ctx.eb, ctx.has_affine, ctx.training_mode = eb, has_affine, noise is not None
        ctx.save_for_backward(z, nz, b, es, med, packed, q, *raw[0], *raw[2])
        # q is differentiable too: HRateHyperprior takes -log q(s) of the side bottleneck outside this function
        # (lossyless/rates.py:640-650); its gradient arrives as g_q and is folded into the same backward launch
        ctx.set_materialize_grads(False)
        return z_hat, q, rates

    @staticmethod
    def backward(ctx, g_zhat, g_q, g_rates):
        eb = ctx.eb
        z, nz, b, es, med, packed, q, *raw = ctx.saved_tensors
        ms, fs = raw[:5], raw[5:]
        rows, C = z.shape
        dev = z.device
        if g_zhat is None and g_q is None and g_rates is None:
            return (None,) * (7 + 14)
        # likelihood gradients arrive through rates (-log q summed inside the kernel) and / or directly through q;
        # eval mode (nz is None): round() passes nothing to z, the parameters still get their gradients
        g_r = g_rates.contiguous().float() if g_rates is not None else None
        g_qc = g_q.contiguous().float() if g_q is not None else None
        gz = torch.empty((rows, C), dtype=torch.float32, device=dev)
        gparam = torch.zeros((N.LL_EB_NPARAM, C), dtype=torch.float32, device=dev)
        gb = torch.zeros(C, dtype=torch.float32, device=dev) if ctx.has_affine else None
        ges = torch.zeros(C, dtype=torch.float32, device=dev) if ctx.has_affine else None
        gh = g_zhat.contiguous().float() if g_zhat is not None else None
        N.call("ll_eb_backward", N.ptr(z), N.dtype_code(z), rows, C, N.ptr(nz), N.ptr(b), N.ptr(es), N.ptr(med), N.ptr(packed),
               _bound(eb), N.ptr(g_r), N.ptr(g_qc), N.ptr(gh), N.ptr(gz), N.ptr(gparam), N.ptr(gb), N.ptr(ges), N.stream())
        # chain rule from the derived parameters back to the raw ones: ONE launch for the 14 tensors
        shapes_m = [(3, 1), (3, 3), (3, 3), (3, 3), (1, 3)]
        grads_m = [torch.empty((C, fo, fi), dtype=torch.float32, device=dev) for fo, fi in shapes_m]
        grads_b = [torch.empty((C, fo, 1), dtype=torch.float32, device=dev) for fo, _ in shapes_m]
        grads_f = [torch.empty((C, 3, 1), dtype=torch.float32, device=dev) for _ in range(4)]
        N.call("ll_eb_unpack_grads", N.ptr(gparam), N.ptr_array(list(ms)), N.ptr_array(list(fs)), C, N.ptr_array(grads_m),
               N.ptr_array(grads_b), N.ptr_array(grads_f), N.stream())
        g_bias = gb if ctx.has_affine else None
        g_scaling = ges * es if ctx.has_affine else None  # d exp(s)/ds
        params_grads = []
        for i in range(5):
            params_grads += [grads_m[i], grads_b[i]]
            if i < 4:
                params_grads.append(grads_f[i])
        return (None, gz.to(z.dtype), None, g_bias, g_scaling, None, None, *params_grads)


class _EBAuxLoss(torch.autograd.Function):
    """quantiles [C, 1, 3] -> scalar auxiliary loss through ll_eb_aux_loss (value + d/d quantiles in one launch)."""

    @staticmethod
    def forward(ctx, eb, quantiles):
        packed, _ = _packed_params(eb, quantiles.device)
        q = quantiles.detach().float().contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=q.device)
        gq = torch.empty_like(q)
        N.call("ll_eb_aux_loss", N.ptr(q), N.ptr(packed), N.ptr(eb.target.detach().float().contiguous()), eb.channels, N.ptr(loss),
               N.ptr(gq), N.stream())
        ctx.save_for_backward(gq)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (gq,) = ctx.saved_tensors
        return None, gq * g


class _GCLikelihood(torch.autograd.Function):
    """(inputs, scales, means) -> (outputs, likelihood, rates) through ll_gc_forward / ll_gc_backward."""

    @staticmethod
    def forward(ctx, gc, inputs, scales, means, noise, want_rates):
        N.require_cuda(inputs, "inputs")
        shape = inputs.shape
        C = shape[-1]
        x = inputs.detach().float().contiguous().reshape(-1, C)
        s = scales.detach().float().expand(shape).contiguous().reshape(-1, C)
        m = means.detach().float().expand(shape).contiguous().reshape(-1, C) if means is not None else None
        nz = noise.detach().float().contiguous().reshape(-1, C) if noise is not None else None
        rows = x.shape[0]
        out = torch.empty_like(x)
        q = torch.empty_like(x)
        rates = torch.empty(rows, dtype=torch.float32, device=x.device) if want_rates else None
        sb = _bound_of(gc.lower_bound_scale)
        lb = _bound_of(gc.likelihood_lower_bound)
        N.call("ll_gc_forward", N.ptr(x), N.ptr(s), N.ptr(m), N.ptr(nz), rows, C, sb, lb, N.ptr(out), N.ptr(q), N.ptr(rates), N.stream())
        ctx.save_for_backward(x, s, m, nz)
        ctx.bounds, ctx.shape, ctx.has_means = (sb, lb), shape, means is not None
        ctx.set_materialize_grads(False)
        return out.reshape(shape), q.reshape(shape), (rates.reshape(shape[:-1]) if want_rates else None)

    @staticmethod
    def backward(ctx, g_out, g_q, g_rates):
        x, s, m, nz = ctx.saved_tensors
        if g_out is None and g_q is None and g_rates is None:
            return (None,) * 6
        rows, C = x.shape
        prep = lambda g, shp: g.float().contiguous().reshape(shp) if g is not None else None  # noqa: E731
        go, gq, gr = prep(g_out, (rows, C)), prep(g_q, (rows, C)), prep(g_rates, (rows,))
        gx, gs = torch.empty_like(x), torch.empty_like(x)
        gm = torch.empty_like(x) if m is not None else None
        N.call("ll_gc_backward", N.ptr(x), N.ptr(s), N.ptr(m), N.ptr(nz), rows, C, ctx.bounds[0], ctx.bounds[1], N.ptr(go), N.ptr(gq),
               N.ptr(gr), N.ptr(gx), N.ptr(gs), N.ptr(gm), N.stream())
        return (None, gx.reshape(ctx.shape), gs.reshape(ctx.shape), gm.reshape(ctx.shape) if gm is not None else None, None, None)


def _bound_of(lower_bound_module):
    """LowerBound.bound as a host float, read back once per buffer version (no `.item()` per call)."""
    buf = lower_bound_module.bound
    key = (buf.data_ptr(), buf._version)
    cache = getattr(lower_bound_module, "_bound_cache", None)
    if cache is None or cache[0] != key:
        cache = lower_bound_module._bound_cache = (key, float(buf.item()))
    return cache[1]


def eb_rate_forward(eb, z, noise, biasing, scaling, want_rates=True, exp_scale=None):
    """Fused ``process_z_in -> EntropyBottleneck -> process_z_out`` (+ per-row rate in nats).

    z [rows, C]; noise [rows, C] float32 (training) or None (eval: dequantise); biasing /
    scaling are the HRateEstimator parameters (None, None: z is already z_in and the returned
    z_hat is in the bottleneck's domain); ``exp_scale`` optionally supplies a precomputed
    exp(scaling) (the codec's CPU-evaluated one).  Returns (z_hat, q, rates); ``rates`` is None when
    ``want_rates`` is False (the per-row reduction and its atomics are skipped)."""
    params = []
    for i in range(5):
        params += [getattr(eb, f"_matrix{i}"), getattr(eb, f"_bias{i}")]
        if i < 4:
            params.append(getattr(eb, f"_factor{i}"))
    return _EBRate.apply(eb, z, noise, biasing, scaling, exp_scale, bool(want_rates), *params)
