"""networks.run_chain scheduling on the CPU (no kernels run: the blocks' forward_act is replaced by recorders):
which producer writes which halo / upsample for its consumer."""
import torch

from munit_b200 import networks as N
from munit_b200.ops import Act


def _record(monkeypatch):
    calls = []

    def conv_fa(self, x, out_pad=0, upsample=1, residual=None, frozen=False):
        calls.append(("conv%d" % self.layer.k, out_pad, upsample, False))
        return x

    def res_fa(self, x, out_pad=0, upsample=1):
        calls.append(("res", out_pad, upsample, False))
        return x

    monkeypatch.setattr(N.Conv2dBlock, "forward_act", conv_fa)
    monkeypatch.setattr(N.ResBlock, "forward_act", res_fa)
    return calls


def _decoder():
    return N.Decoder(2, 4, 256, 3, res_norm="adain", activ="relu", pad_type="reflect")


def test_default_schedule_producer_writes_upsample_and_halo(monkeypatch):
    calls = _record(monkeypatch)
    dec = _decoder()
    N.run_chain(list(dec.model), Act(torch.zeros(1), 1), 0)
    assert calls == [("res", 1, 1, False)] * 3 + [("res", 2, 2, False), ("conv5", 2, 2, False), ("conv5", 3, 1, False),
                                                  ("conv7", 0, 1, False)]
    # content encoder: 7x7 -> 4x4 s2 -> 4x4 s2 -> 4 residual blocks, final halo chosen by the caller
    calls.clear()
    enc = N.ContentEncoder(2, 4, 3, 64, "in", "relu", pad_type="reflect")
    N.run_chain(list(enc.model), torch.zeros(1, 3, 8, 8), 1)
    assert calls == [("conv7", 1, 1, False), ("conv4", 1, 1, False), ("conv4", 1, 1, False)] + [("res", 1, 1, False)] * 4
