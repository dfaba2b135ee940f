"""GPU suite, SURVEY.md 8f row N4 (first slice): the Joints head's training step on the device -- forward with saved activations,
the loss of joints.py:54-75 and the full backward pass -- against
  (1) the fixture the LIVE reference produced (its own `Joints.shared_step` + `loss.backward()`, eval mode and with a fixed dropout
      mask: tests/golden/train_joints_step.npz, oracle/make_golden_train.py),
  (2) the oracle (oracle/train_port.py: the same torch modules under autograd) on a larger ragged batch, every gradient tensor."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _module(seeded_state_dict):
    import mobileposer_b200 as mp
    m = mp.Joints()
    m.load_state_dict({k[len('joints.'):]: v for k, v in seeded_state_dict.items() if k.startswith('joints.')})
    return m.to(DEV)


def _rel(a, b):
    return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize('tag', ['eval', 'mask'])
def test_joints_training_step_against_the_live_reference(tag):
    import mobileposer_b200 as mp
    from mobileposer_b200.training import joints_shared_step
    g = load_golden('train_joints_step')
    torch.manual_seed(0)                 # the fixture's module: torch.manual_seed(0); Joints() -- the same constructor order here
    mod = mp.Joints().to(DEV)
    mask = g['mask'].to(DEV) if tag == 'mask' else None
    loss, grads, _ = joints_shared_step(mod, g['imu'].to(DEV), g['lengths'].tolist(), g['target'].to(DEV), mask)
    assert abs(loss.item() - g[f'{tag}_loss'].item()) <= 1e-6 * abs(g[f'{tag}_loss'].item()) + 1e-9
    worst = 0.0
    for name, gr in grads.items():
        short = name[len('joints.'):]
        ref = g[f'{tag}_grad.{short}']
        got = gr.cpu() if gr.numel() <= 40000 else gr[::7, ::5].cpu()
        worst = max(worst, _rel(got, ref))
        assert _rel(got, ref) < 2e-4, (name, _rel(got, ref))
        assert abs(gr.norm().item() - g[f'{tag}_norm.{short}'].item()) <= 2e-4 * g[f'{tag}_norm.{short}'].item(), name
    print(f'[train] {tag}: loss {loss.item():.6f}, worst relative gradient error over 20 tensors {worst:.2e}')


def test_joints_training_step_against_the_oracle_on_a_ragged_batch(seeded_state_dict):
    from mobileposer_b200.synthetic import synthetic_imu_batch
    from mobileposer_b200.training import dropout_mask, joints_shared_step
    from oracle.train_port import joints_shared_step as oracle_step
    B, T = 19, 37                                   # three CTAs of 8 sequences, the last one partial
    gen = torch.Generator().manual_seed(5)
    lens = [int(v) for v in torch.randint(1, T + 1, (B,), generator=gen)]
    lens[4] = T
    imu = synthetic_imu_batch(list(range(400, 400 + B)), T)
    for b, L in enumerate(lens):
        imu[b, L:] = 0
    target = torch.randn(B, T, 72, generator=gen) * 0.3
    mask = dropout_mask((B, T, 256), generator=gen)
    sd = {k: v for k, v in seeded_state_dict.items() if k.startswith('joints.')}
    o_loss, o_grads, o_pred = oracle_step(sd, imu, lens, target, mask, prefix='joints.joints.')
    loss, grads, pred = joints_shared_step(_module(seeded_state_dict), imu.to(DEV), lens, target.to(DEV), mask.to(DEV))
    assert (pred.cpu() - o_pred).abs().max() < 1e-5
    assert abs(loss.item() - o_loss.item()) <= 1e-6 * abs(o_loss.item())
    for k, ref in o_grads.items():
        assert _rel(grads['joints.' + k], ref) < 2e-4, (k, _rel(grads['joints.' + k], ref))


@pytest.mark.parametrize('name', ['foot', 'vel'])
def test_footcontact_and_velocity_training_steps_against_the_live_reference(name):
    """FootContact.shared_step (BCE with logits, H = 64 bidirectional) and Velocity.shared_step (windowed MSE, H = 256
    unidirectional) + backward: loss and every gradient tensor against the live reference's (tests/golden/train_heads_step.npz;
    the reference's in-step noise is reproduced in the fixture's input)."""
    import mobileposer_b200 as mp
    from mobileposer_b200.training import footcontact_shared_step, velocity_shared_step
    g = load_golden('train_heads_step')
    torch.manual_seed(0)
    if name == 'foot':
        mod, step, target, prefix = mp.FootContact().to(DEV), footcontact_shared_step, g['foot_contacts'], 'footcontact.'
    else:
        mod, step, target, prefix = mp.Velocity().to(DEV), velocity_shared_step, g['vels'].view(3, 20, 72), 'vel.'
    loss, grads, _ = step(mod, g[f'{name}_input'].to(DEV), g['lengths'].tolist(), target.to(DEV))
    assert abs(loss.item() - g[f'{name}_loss'].item()) <= 2e-6 * abs(g[f'{name}_loss'].item()) + 1e-9
    worst = 0.0
    for pname, gr in grads.items():
        short = pname[len(prefix):]
        ref = g[f'{name}_grad.{short}']
        got = gr.cpu() if gr.numel() <= 40000 else gr[::7, ::5].cpu()
        worst = max(worst, _rel(got, ref))
        assert _rel(got, ref) < 2e-4, (pname, _rel(got, ref))
        assert abs(gr.norm().item() - g[f'{name}_norm.{short}'].item()) <= 2e-4 * g[f'{name}_norm.{short}'].item(), pname
    print(f'[train] {name}: loss {loss.item():.6f}, worst relative gradient error over {len(grads)} tensors {worst:.2e}')


def test_poser_training_step_against_the_live_reference():
    """Poser.shared_step (poser.py:65-98: MSE + jerk L1 + joint-position loss through _reduced_global_to_full and the zero-pose
    forward kinematics) + backward: the loss kernel carries the Gram-Schmidt and kinematic-tree adjoints; loss and every gradient
    tensor against the live reference's autograd."""
    import mobileposer_b200 as mp
    from mobileposer_b200.training import poser_shared_step
    g = load_golden('train_heads_step')
    torch.manual_seed(0)
    mod = mp.Poser().to(DEV)
    loss, grads, _ = poser_shared_step(mod, g['pose_input'].to(DEV), g['lengths'].tolist(), g['poses'].to(DEV), g['joints_gt'].to(DEV))
    assert abs(loss.item() - g['pose_loss'].item()) <= 2e-6 * abs(g['pose_loss'].item())
    worst = 0.0
    for pname, gr in grads.items():
        short = pname[len('pose.'):]
        ref = g[f'pose_grad.{short}']
        got = gr.cpu() if gr.numel() <= 40000 else gr[::7, ::5].cpu()
        worst = max(worst, _rel(got, ref))
        assert _rel(got, ref) < 2e-4, (pname, _rel(got, ref))
        assert abs(gr.norm().item() - g[f'pose_norm.{short}'].item()) <= 2e-4 * g[f'pose_norm.{short}'].item(), pname
    print(f'[train] pose: loss {loss.item():.6f}, worst relative gradient error over {len(grads)} tensors {worst:.2e}')


@pytest.mark.parametrize('tag,clip', [('clip1', 1.0), ('clip005', 0.005)])
def test_overfit_loop_against_the_live_reference(tag, clip):
    """HeadTrainer = the loop Lightning runs for overfit.py:41-56 (zero_grad, shared_step, backward, clip_grad_norm_, the AdamW of
    joints.py:113-114): 6 steps on one fixed batch with a fixed dropout mask; the loss of every step, the gradient norms and the
    final parameters against the live reference's module + torch.optim.AdamW (tests/golden/train_overfit_joints.npz)."""
    import mobileposer_b200 as mp
    from mobileposer_b200.training import HeadTrainer
    g = load_golden('train_overfit_joints')
    torch.manual_seed(0)
    mod = mp.Joints().to(DEV)
    tr = HeadTrainer(mod, gradient_clip_val=clip)
    imu, lens, target, mask = g['imu'].to(DEV), g['lengths'].tolist(), g['target'].to(DEV), g['mask'].to(DEV)
    losses, norms = [], []
    for _ in range(6):
        losses.append(tr.training_step(imu, lens, target, mask=mask).item())
        norms.append(tr.grad_norm())
    ref_l, ref_n = g[f'{tag}_losses'].double(), g[f'{tag}_grad_norms'].double()
    worst_l = max(abs(a - b.item()) / abs(b.item()) for a, b in zip(losses, ref_l))
    worst_n = max(abs(a - b.item()) / abs(b.item()) for a, b in zip(norms, ref_n))
    assert worst_l < 2e-5, (losses, ref_l)
    assert worst_n < 2e-3, (norms, ref_n)
    worst_p = 0.0
    for name, p in mod.joints.named_parameters():
        ref = g[f'{tag}_param.{name}']
        got = p.detach().cpu() if p.numel() <= 40000 else p.detach()[::7, ::5].cpu()
        worst_p = max(worst_p, (got - ref).abs().max().item())
        assert (got - ref).abs().max().item() < 2e-5, (name, (got - ref).abs().max().item())      # six steps of lr = 1e-3 move a weight by <= 6e-3
        assert abs(p.norm().item() - g[f'{tag}_pnorm.{name}'].item()) <= 1e-5 * g[f'{tag}_pnorm.{name}'].item(), name
    # the trained values are what the inference path sees (the parameters are views of the flat buffer; versions were bumped)
    y, _, _ = mod.joints(imu, lens)
    assert torch.isfinite(y).all()
    print(f'[train] overfit loop {tag}: losses {[round(v, 6) for v in losses]}, worst relative loss error {worst_l:.1e}, '
          f'gradient-norm error {worst_n:.1e}, worst final-parameter error {worst_p:.1e}')


def test_adamw_kernel_against_torch_on_a_ragged_buffer():
    """mp_adamw_step / mp_grad_sq_norm on a flat buffer whose length is not a multiple of 4, 3 steps, against torch.optim.AdamW +
    clip_grad_norm_ on the same numbers."""
    import ctypes as C
    from mobileposer_b200 import _cabi
    lib = _cabi.lib()
    n = 4099
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=gen)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3)
    p = torch.zeros(n + 1, device=DEV)[:n]
    p.copy_(p0)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sq = torch.zeros((), device=DEV, dtype=torch.float64)
    s = torch.cuda.current_stream().cuda_stream
    for step in range(1, 4):
        gr = torch.randn(n, generator=gen) * (3.0 if step == 2 else 0.01)
        ref.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        gd = gr.to(DEV)
        sq.zero_()
        _cabi.check(lib.mp_grad_sq_norm(gd.data_ptr(), n, sq.data_ptr(), s))
        assert abs(sq.sqrt().item() - gr.double().norm().item()) < 1e-9 * gr.double().norm().item() + 1e-12
        _cabi.check(lib.mp_adamw_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step,
                                      sq.data_ptr(), 1.0, 1.0, s))
        assert (p.cpu() - ref.detach()).abs().max().item() < 2e-6, step


@pytest.mark.parametrize('kind', ['velocity', 'footcontact', 'poser'])
def test_head_trainer_on_the_other_heads_against_the_oracle_loop(kind):
    """HeadTrainer (clip + AdamW on the device) for the unidirectional H = 256, the H = 64 and the Poser head: 3 optimisation steps on
    one fixed batch against oracle/train_port.py:overfit_loop (the loop pinned to the live reference for the Joints head)."""
    import mobileposer_b200 as mp
    from mobileposer_b200.config import joint_set
    from mobileposer_b200.synthetic import synthetic_imu_batch
    from mobileposer_b200.training import HeadTrainer
    from oracle.train_port import overfit_loop
    B, T = 5, 24
    gen = torch.Generator().manual_seed(17)
    lens = [24, 9, 24, 17, 3]
    imu = synthetic_imu_batch(list(range(40, 40 + B)), T)
    for b, L in enumerate(lens):
        imu[b, L:] = 0
    x = torch.cat((torch.randn(B, T, 72, generator=gen) * 0.3, imu), -1)
    torch.manual_seed(0)
    if kind == 'velocity':
        mod, prefix, attr = mp.Velocity(), 'vel.', 'vel'
        args = (x, lens, torch.randn(B, T, 72, generator=gen) * 0.5)
        o_target = args[2]
    elif kind == 'footcontact':
        mod, prefix, attr = mp.FootContact(), 'footcontact.', 'footcontact'
        args = (x, lens, (torch.rand(B, T, 2, generator=gen) > 0.5).float())
        o_target = args[2]
    else:
        mod, prefix, attr = mp.Poser(), 'pose.', 'pose'
        eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
        poses = eye6 + 0.3 * torch.randn(B, T, 144, generator=gen)
        joints_t = torch.randn(B, T, 72, generator=gen) * 0.3
        args = (x, lens, poses, joints_t)
        o_target = torch.cat((poses.view(B, T, 24, 6)[:, :, joint_set.reduced].reshape(B, T, 96), joints_t), -1)
    H = getattr(mod, attr).n_hidden
    mask = (torch.rand(B, T, H, generator=gen) >= 0.4).float() / 0.6
    sd = {prefix + k: v.detach().clone() for k, v in getattr(mod, attr).state_dict().items()}
    o_losses, o_final = overfit_loop(sd, x, lens, o_target, mask, 3, prefix=prefix, kind=kind, gradient_clip_val=0.5)
    tr = HeadTrainer(mod.to(DEV), gradient_clip_val=0.5)
    dargs = tuple(a.to(DEV) if torch.is_tensor(a) else a for a in args)
    losses = [tr.training_step(*dargs, mask=mask.to(DEV)).item() for _ in range(3)]
    worst_l = max(abs(a - b.item()) / abs(b.item()) for a, b in zip(losses, o_losses))
    assert worst_l < 5e-5, (losses, o_losses)
    worst_p = 0.0
    for name, p in getattr(mod, attr).named_parameters():
        worst_p = max(worst_p, (p.detach().cpu() - o_final[name]).abs().max().item())
    # three steps of lr = 1e-3 move a weight by <= 3e-3; AdamW's m / (sqrt(v) + eps) amplifies gradient noise where |g| ~ eps
    assert worst_p < 1e-4, worst_p
    print(f'[train] HeadTrainer {kind}: losses {[round(v, 6) for v in losses]}, worst relative loss error {worst_l:.1e}, worst parameter error {worst_p:.1e}')
