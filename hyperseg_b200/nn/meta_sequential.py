"""Sequential container that routes generated weights to its dynamic children.

Mirror of the reference's ``MetaSequential`` (hyperseg/models/layers/meta_sequential.py:10-40): same
constructor, same ``hyper_params`` / ``_ranges`` bookkeeping, same ``forward(x, w)`` contract where ``w``
is either one tensor sliced along dim 1 by ``_ranges`` or a list with one entry per dynamic child.

Two B200-specific differences, both invisible to callers:
  * a dynamic child followed by an eval-mode ``BatchNorm2d`` and an optional ``ReLU``/``ReLU6`` hands
    those to the child's kernel epilogue (they stay registered as modules, so state_dict keys and
    ``load_state_dict`` are unchanged) -- reference blocks of that shape are built by
    make_hyper_patch_conv2d_block (hyperseg_v1_0.py:748-760) and make_meta_patch_conv2d_block
    (meta_patch.py:228-257);
  * children that declare ``accepts_strided_weights`` receive the channel slice as a view instead of
    the reference's ``.contiguous()`` copy (meta_sequential.py:35).
"""
import torch.nn as nn


def _activation_code(module):
    if isinstance(module, nn.ReLU6):
        return "relu6"
    if isinstance(module, nn.ReLU):
        return "relu"
    return None


class MetaSequential(nn.Sequential):
    accepts_strided_weights = True     # it only slices further

    def __init__(self, *args):
        super().__init__(*args)
        self.hyper_params = 0
        self._ranges = [0]
        for child in self:
            self.hyper_params += getattr(child, "hyper_params", 0)
            self._ranges.append(self.hyper_params)

    def _epilogue_for(self, index, modules):
        """(norm, act_code, consumed) for the static modules that can be fused after modules[index]."""
        child = modules[index]
        if not getattr(child, "supports_fused_epilogue", False):
            return None, None, 0
        consumed, norm, act = 0, None, None
        nxt = index + 1
        if nxt < len(modules) and isinstance(modules[nxt], nn.BatchNorm2d) and not modules[nxt].training \
                and modules[nxt].running_mean is not None:
            norm = modules[nxt]
            consumed += 1
            nxt += 1
        if nxt < len(modules) and _activation_code(modules[nxt]) is not None:
            act = _activation_code(modules[nxt])
            consumed += 1
        return norm, act, consumed

    def forward(self, x, w):
        modules = list(self)
        taken = 0      # index into a list-valued w
        i = 0
        while i < len(modules):
            child = modules[i]
            lo, hi = self._ranges[i], self._ranges[i + 1]
            if lo < hi:
                if isinstance(w, (list, tuple)):
                    cw = w[taken]
                else:
                    cw = w[:, lo:hi]
                    if not getattr(child, "accepts_strided_weights", False):
                        cw = cw.contiguous()
                taken += 1
                norm, act, consumed = self._epilogue_for(i, modules)
                if consumed:
                    x = child(x, cw, fused_norm=norm, fused_act=act)
                    i += consumed
                else:
                    x = child(x, cw)
            else:
                x = child(x)
            i += 1
        return x
