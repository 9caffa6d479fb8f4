"""GPU: the replayed training step (engine.TrainStep) keeps doing what the eager loop of train.py:187-214 does.

Regression tests for two defects of the first round: a StepLR change never reached the captured Adam
launch, and in multi-GPU ("split") graph mode the packed weight gradients were dropped from the second replay on because
the host-side `pending` flag is only set by the Python backward.  Plus: optimiser / scheduler state interchange with
torch.optim.Adam / StepLR (the reference's checkpoint schema, train.py:123-140, 285-306)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _setup(features=8, lr=1e-3, use_graph=True, warmup=2):
    from b200seg.engine import TrainStep
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.optim import FusedAdam
    from b200seg.utils.loss_function import DiceCELoss
    from oracle import unet3d as ounet
    sd = ounet.init_state_dict(1, 2, features, seed=3)
    net = UNet3D(1, 2, features).to(DEV)
    net.load_state_dict(sd)
    net.train()
    opt = FusedAdam(net.parameters(), lr=lr)
    return net, opt, TrainStep(net, DiceCELoss(2), opt, use_graph=use_graph, warmup=warmup)


def _data(n, seed=5):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [(torch.randn(2, 1, 32, 32, 32, device=DEV, generator=g),
             (torch.rand(2, 32, 32, 32, device=DEV, generator=g) > 0.8).to(torch.uint8)) for _ in range(n)]


def test_lr_change_reaches_the_replayed_graph():
    net, opt, step = _setup()
    data = _data(6)
    for x, y in data[:4]:
        step(x, y)
    assert step.graph is not None
    w = net.decoder1[3].weight
    before = w.detach().clone()
    opt.lr = 0.0                      # what StepLR does between epochs (train.py:259-261), taken to the extreme
    step(*data[4])
    torch.cuda.synchronize()
    assert torch.equal(w.detach(), before), "the replayed Adam launch ignored the new learning rate"
    opt.lr = 1e-3
    step(*data[5])
    torch.cuda.synchronize()
    assert not torch.equal(w.detach(), before)
    # and the trajectory under a decaying schedule equals the eager one
    from b200seg.train import StepLR
    finals = []
    for use_graph in (False, True):
        net, opt, step = _setup(use_graph=use_graph)
        sched = StepLR(opt, 2, 0.1)
        for x, y in data:
            step(x, y)
            sched.step()
        torch.cuda.synchronize()
        assert abs(opt.lr - 1e-3 * 0.1 ** 3) < 1e-12
        finals.append(net.decoder1[3].weight.detach().clone())
    assert rel(finals[1], finals[0]) < 2e-2, rel(finals[1], finals[0])   # atomics order + Adam's sign-like early steps


def test_split_mode_replay_keeps_the_weight_gradients(monkeypatch):
    """Multi-GPU graph mode on one GPU: forward + backward (+ the weight-gradient transpose) are captured, the gradient
    exchange and Adam run after every replay.  The conv weights must keep moving like in the eager loop."""
    from b200seg import engine
    data = _data(8, seed=9)
    finals = []
    for use_graph in (False, True):
        net, opt, step = _setup(features=16, use_graph=use_graph)
        if use_graph:
            monkeypatch.setattr(engine.parallel, "is_parallel", lambda group=None: True)
            monkeypatch.setattr(engine.parallel, "graph_safe", lambda group=None: True)
        for x, y in data:
            step(x, y)
        torch.cuda.synchronize()
        if use_graph:
            assert step.graph is not None and step._split
            monkeypatch.undo()
        finals.append({k: v.detach().clone() for k, v in net.state_dict().items()})
    for k in ("decoder1.dec1conv2.weight", "encoder1.enc1conv1.weight", "upconv1.weight", "conv.weight"):
        e = rel(finals[1][k], finals[0][k])
        assert e < 2e-2, (k, e)


def test_optimizer_and_scheduler_state_interchange_with_torch():
    from b200seg.optim import FusedAdam
    from b200seg.train import StepLR
    g = torch.Generator(device=DEV).manual_seed(1)
    shapes = [(16, 8, 3, 3, 3), (16,), (8, 16, 2, 2, 2), (5,)]
    params = [torch.nn.Parameter(torch.randn(*s, device=DEV, generator=g) * 0.1) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in params]
    topt = torch.optim.Adam(ref, lr=2e-3, betas=(0.8, 0.99), eps=1e-7, weight_decay=0.01)
    tsch = torch.optim.lr_scheduler.StepLR(topt, step_size=2, gamma=0.5)
    grads = [[torch.randn(*s, device=DEV, generator=g) for s in shapes] for _ in range(5)]
    for it in range(3):
        for r, gr in zip(ref, grads[it]):
            r.grad = gr.clone()
        topt.step()
        tsch.step()
    # a torch (= reference) checkpoint resumes here ...
    for p, r in zip(params, ref):
        p.data.copy_(r.data)
    opt = FusedAdam(params, lr=1.0)
    sch = StepLR(opt, 7, 0.9)
    opt.load_state_dict(topt.state_dict())
    sch.load_state_dict(tsch.state_dict())
    assert opt.betas == (0.8, 0.99) and opt.eps == 1e-7 and opt.weight_decay == 0.01 and opt.step_count == 3
    assert abs(opt.lr - tsch.get_last_lr()[0]) < 1e-12 and sch.step_size == 2
    for it in (3, 4):
        opt.zero_grad()
        for p, r, gr in zip(params, ref, grads[it]):
            p.grad.copy_(gr)
            r.grad = gr.clone()
        opt.step()
        topt.step()
        sch.step()
        tsch.step()
        assert abs(opt.lr - tsch.get_last_lr()[0]) < 1e-12
    for p, r in zip(params, ref):
        assert rel(p.detach(), r.detach()) < 1e-5
    # ... and a checkpoint written here resumes under torch.optim.Adam / StepLR
    t2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in params], lr=1.0)
    t2.load_state_dict(opt.state_dict())
    s2 = torch.optim.lr_scheduler.StepLR(t2, step_size=1, gamma=1.0)
    s2.load_state_dict(sch.state_dict())
    assert s2.step_size == 2 and s2.last_epoch == sch.last_epoch
    st, rt = t2.state_dict(), topt.state_dict()
    for i in range(len(params)):
        assert rel(st["state"][i]["exp_avg"], rt["state"][i]["exp_avg"]) < 1e-5
        assert float(st["state"][i]["step"]) == float(rt["state"][i]["step"]) == 5
    # wrong shapes are refused
    bad = topt.state_dict()
    bad["state"][0]["exp_avg"] = torch.zeros(3, 3)
    with pytest.raises(ValueError):
        opt.load_state_dict(bad)
