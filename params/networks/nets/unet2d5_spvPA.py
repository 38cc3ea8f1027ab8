"""UNet2d5_spvPA: 2.5D U-Net with spatial attention gates — same constructor, attributes, nesting
(hence ``state_dict`` keys) and ``forward`` contract as the reference
(/root/reference/params/networks/nets/unet2d5_spvPA.py:24-209), executed on B200 by the native
sm_100a kernels.

forward(x [B,1,X,Y,Z] float32) -> (logits [B,out,X,Y,Z], att_maps: list of [B,1,x,y,z], coarsest
first; [] when attention_module=False).

  * CUDA + eval : the fused launch plan of vs_seg_b200.engine.UNetEvalPlan (cached per input
    shape and invalidated when any parameter/buffer changes).
  * CUDA + train: native training kernels (vs_seg_b200.training) — raises if unavailable; there
    is no eager-torch CUDA fallback.
  * CPU tensors : the torch containers run as ordinary modules (host plumbing, BASELINE config 1).
"""
import torch
import torch.nn as nn

from params.networks.blocks.attentionblock import AttentionBlock1, AttentionBlock2
from params.networks.blocks.convolutions import Convolution, ResidualUnit
from vs_seg_b200.compat import Act, Norm, SkipConnection


class UNet2d5_spvPA(nn.Module):
    def __init__(
        self,
        dimensions,
        in_channels,
        out_channels,
        channels,
        strides,
        kernel_sizes,
        sample_kernel_sizes,
        num_res_units=0,
        act=Act.PRELU,
        norm=Norm.INSTANCE,
        dropout=0,
        attention_module=True,
    ):
        super().__init__()
        assert len(channels) == len(kernel_sizes) == (len(strides)) + 1 == len(sample_kernel_sizes) + 1
        self.dimensions = dimensions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.channels = channels
        self.strides = strides
        self.kernel_sizes = kernel_sizes
        self.sample_kernel_sizes = sample_kernel_sizes
        self.num_res_units = num_res_units
        self.act = act
        self.norm = norm
        self.dropout = dropout
        self.attention_module = attention_module
        self.att_maps = []
        self._plans = {}

        def build_level(inc, outc, lvl, is_top):
            # one resolution level: encoder unit, [downsample -> deeper levels -> upsample] wrapped in a
            # skip connection, decoder unit; recursion bottoms out in the bottom layer
            c, s, k, sk = channels[lvl], strides[lvl], kernel_sizes[lvl], sample_kernel_sizes[lvl]
            down = self._get_down_layer(in_channels=inc, out_channels=c, kernel_size=k)
            downsample = self._get_downsample_layer(in_channels=c, out_channels=c, strides=s, kernel_size=sk)
            if lvl + 2 < len(channels):
                sub = build_level(c, channels[lvl + 1], lvl + 1, False)
            else:
                sub = self._get_bottom_layer(in_channels=c, out_channels=channels[lvl + 1],
                                             kernel_size=kernel_sizes[lvl + 1])
            upsample = self._get_upsample_layer(in_channels=channels[lvl + 1], out_channels=c, strides=s,
                                                up_kernel_size=sk)
            up = self._get_up_layer(in_channels=2 * c, out_channels=outc, kernel_size=k, is_top=is_top)
            return nn.Sequential(down, SkipConnection(nn.Sequential(downsample, sub, upsample)), up)

        self.model = build_level(in_channels, out_channels, 0, True)

        if self.attention_module:
            for layer in self.model.modules():
                if type(layer) == AttentionBlock1:
                    layer.register_forward_hook(self.hook_save_attention_map)

    # ---- layer factories (names kept from the reference) -----------------------------------
    def hook_save_attention_map(self, module, inp, outp):
        if len(self.att_maps) == len(self.channels):
            self.att_maps = []
        self.att_maps.append(outp[0])

    def _get_att_layer(self, in_channels, out_channels, kernel_size):
        att1 = AttentionBlock1(self.dimensions, in_channels, out_channels, kernel_size, norm=None, dropout=self.dropout)
        att2 = AttentionBlock2(self.dimensions, in_channels, out_channels, kernel_size, norm=None, dropout=self.dropout)
        return nn.Sequential(att1, att2)

    def _get_down_layer(self, in_channels, out_channels, kernel_size):
        if self.num_res_units > 0:
            return ResidualUnit(self.dimensions, in_channels, out_channels, strides=1, kernel_size=kernel_size,
                                subunits=self.num_res_units, act=self.act, norm=self.norm, dropout=self.dropout)
        return Convolution(self.dimensions, in_channels, out_channels, strides=1, kernel_size=kernel_size,
                           act=self.act, norm=self.norm, dropout=self.dropout)

    def _get_downsample_layer(self, in_channels, out_channels, strides, kernel_size):
        return Convolution(self.dimensions, in_channels, out_channels, strides, kernel_size, self.act, self.norm,
                           self.dropout, is_transposed=False)

    def _get_bottom_layer(self, in_channels, out_channels, kernel_size):
        conv = self._get_down_layer(in_channels, out_channels, kernel_size)
        if self.attention_module:
            return nn.Sequential(self._get_att_layer(in_channels, in_channels, kernel_size), conv)
        return conv

    def _get_upsample_layer(self, in_channels, out_channels, strides, up_kernel_size):
        return Convolution(self.dimensions, in_channels, out_channels, strides, up_kernel_size, self.act, self.norm,
                           self.dropout, is_transposed=True)

    def _get_up_layer(self, in_channels, out_channels, kernel_size, is_top):
        att_layer = self._get_att_layer(in_channels, in_channels, kernel_size) if self.attention_module else None
        ru = None
        if self.num_res_units > 0:
            ru = ResidualUnit(self.dimensions, in_channels, out_channels, strides=1, kernel_size=kernel_size,
                              subunits=1, act=self.act, norm=self.norm, dropout=self.dropout, last_conv_only=is_top)
        if att_layer is not None and ru is not None:
            return nn.Sequential(att_layer, ru)
        if att_layer is not None:
            return att_layer
        if ru is not None:
            return ru
        return nn.Identity

    # ---- native execution ------------------------------------------------------------------
    def _weights_version(self):
        return tuple(t._version for t in self.state_dict(keep_vars=True).values())

    def _plan_supported(self):
        name = self.norm if isinstance(self.norm, str) else self.norm[0]
        act = self.act if isinstance(self.act, str) else self.act[0]
        return (self.dimensions == 3 and self.num_res_units == 2 and str(name).upper() == "BATCH"
                and str(act).upper() == "PRELU" and self.in_channels == 1 and self.out_channels in (1, 2))

    def eval_plan(self, patch_size, batch=1, device=None, window_levels=0, slot=0, atomic_out=False):
        """The cached fused launch plan for eval-mode inference on [batch,1,*patch_size]
        (window_levels: see vs_seg_b200.engine.UNetEvalPlan; slot: a second plan with its own activation buffers,
        used when two window groups run concurrently on two streams; atomic_out: the caller blends with atomics, so
        the last launch takes all windows of the group at once)."""
        from vs_seg_b200.engine import UNetEvalPlan, batch_first_enabled
        if not self._plan_supported():
            raise NotImplementedError("the native plan covers the reference configuration "
                                      "(3-D, num_res_units=2, BatchNorm, PReLU, 1 input channel)")
        device = torch.device(device) if device is not None else next(self.parameters()).device
        key = (tuple(int(v) for v in patch_size), int(batch), str(device), int(window_levels), int(slot),
               batch_first_enabled(), bool(atomic_out))
        ver = self._weights_version()
        hit = self._plans.get(key)
        if hit is None or hit[0] != ver:
            if len(self._plans) >= 8:
                self._plans.clear()
            plan = UNetEvalPlan(self.state_dict(), key[0], batch=batch, device=device, attention=self.attention_module,
                                window_levels=window_levels, atomic_out=atomic_out,
                                channels=self.channels, strides=self.strides, kernel_sizes=self.kernel_sizes,
                                sample_kernel_sizes=self.sample_kernel_sizes, in_channels=self.in_channels,
                                out_channels=self.out_channels)
            self._plans[key] = hit = (ver, plan)
        return hit[1]

    def forward(self, x):
        if x.is_cuda:
            if self.training:
                from vs_seg_b200.training import unet_train_forward
                logits, self.att_maps = unet_train_forward(self, x)
                return logits, self.att_maps
            plan = self.eval_plan(x.shape[2:], batch=x.shape[0], device=x.device)
            logits, atts = plan.forward(x.float().contiguous())
            self.att_maps = [a.clone() for a in atts]
            return logits, self.att_maps
        x = self.model(x)
        return x, self.att_maps


Unet2d5_spvPA = unet2d5_spvPA = UNet2d5_spvPA
