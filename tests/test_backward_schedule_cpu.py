"""Host logic of the two-stream backward schedule (onet_b200/model.py: _wgrad / _flush_side / _join_side / _block_done)
without a GPU: fake streams and a recording `call`.  Invariants:
  * a deferred weight gradient is issued on the side stream, after the side stream waited for the main stream;
  * its operands stay referenced until the main stream has waited for the side stream;
  * a block's all-reduce hook fires only after a join that follows the flush which issued the block's last weight gradient;
  * with the side stream disabled everything is issued at once on the main stream, hooks fire immediately."""
import onet_b200.model as M


class _FakeStream:
    def __init__(self, name, log):
        self.name, self.log, self.cuda_stream = name, log, hash(name) & 0xffff

    def wait_stream(self, other):
        self.log.append(("wait", self.name, other.name))


def _engine(log, side=True):
    eng = object.__new__(M._Engine)
    eng._main = _FakeStream("main", log)
    eng._side = _FakeStream("side", log) if side else None
    eng.stream = eng._main.cuda_stream
    eng._deferred, eng._inflight, eng._hooks_deferred, eng._hooks_ready = [], [], [], []
    eng._after_block = lambda unet, block: log.append(("hook", block))
    return eng


def test_two_stream_schedule_order(monkeypatch):
    log = []
    monkeypatch.setattr(M, "call", lambda name, *args: log.append(("call", name, args[-1])))
    eng = _engine(log)
    side, main = eng._side.cuda_stream, eng._main.cuda_stream
    keep = object()
    eng._wgrad((keep,), "onet_conv3x3_wgrad", 1, 2)            # layer l: deferred, nothing issued yet
    eng._block_done("unet", "up4")
    assert log == [] and len(eng._deferred) == 1
    eng._join_side()                                            # a join before the flush must not release the hook
    assert log == []
    eng._flush_side()                                           # just before bn_bwd(l-1) goes to the main stream
    assert log == [("wait", "side", "main"), ("call", "onet_conv3x3_wgrad", side)]
    assert eng._inflight == [(keep,)] and eng._deferred == []
    eng._wgrad((), "onet_convT2x2_wgrad", 3)                    # deferred behind the flush: belongs to the NEXT join
    eng._block_done("unet", "up3")
    eng._join_side()                                            # before dgrad(l-1)
    assert log[2:] == [("wait", "main", "side"), ("hook", "up4")]
    assert eng._inflight == [] and len(eng._deferred) == 1 and eng._hooks_deferred == [("unet", "up3")]
    eng._flush_side()
    eng._join_side()
    assert log[4:] == [("wait", "side", "main"), ("call", "onet_convT2x2_wgrad", side), ("wait", "main", "side"), ("hook", "up3")]
    eng._flush_side()                                           # nothing pending: no events, no hooks
    eng._join_side()
    assert len(log) == 8
    assert all(ev[2] != main for ev in log if ev[0] == "call")


def test_serial_schedule_when_side_stream_is_off(monkeypatch):
    log = []
    monkeypatch.setattr(M, "call", lambda name, *args: log.append(("call", name, args[-1])))
    eng = _engine(log, side=False)
    eng._wgrad((), "onet_conv3x3_wgrad", 1)
    eng._block_done("unet", "up4")
    eng._flush_side()
    eng._join_side()
    assert log == [("call", "onet_conv3x3_wgrad", eng.stream), ("hook", "up4")]


def test_cosine_warm_restarts_matches_torch_scheduler():
    """onet_b200.trainer.cosine_warm_restarts_lr against torch's CosineAnnealingWarmRestarts stepped once per epoch, with the
    ZY-3 script's settings (Train_Onet_on_zy3_20240606.py:88-89) and a short-period variant that crosses several restarts."""
    import torch
    from onet_b200.trainer import cosine_warm_restarts_lr
    for base_lr, T_0, T_mult, eta_min, epochs in ((1e-4, 300, 2, 1e-6, 1000), (3e-4, 7, 2, 1e-5, 120), (1e-3, 5, 1, 0.0, 23)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=base_lr)
        sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=T_0, T_mult=T_mult, eta_min=eta_min)
        for epoch in range(epochs):
            want = opt.param_groups[0]["lr"]
            got = cosine_warm_restarts_lr(epoch, base_lr, T_0, T_mult, eta_min)
            assert abs(got - want) <= 1e-12 + 1e-9 * want, (epoch, got, want)
            opt.step()
            sch.step()
