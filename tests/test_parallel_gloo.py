"""world_size-2 gloo test (CPU) of the data-parallel host logic: disjoint image assignment from a rank-shared
permutation, flat-gradient all-reduce, densification-state reduction (sum / max)."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from gssdf_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs = [parallel.image_for_rank(s, rank, world, 10, seed=3) for s in range(12)]
    flat = torch.full((1000,), float(rank + 1))
    parallel.allreduce_flat_grad(flat, world, average=True)
    st = dict(grad2d=torch.arange(5.0) * (rank + 1), count=torch.ones(5), vis=torch.tensor([0.1, 0.9]) if rank == 0 else torch.tensor([0.5, 0.2]),
              radii=torch.tensor([float(rank)]))
    parallel.allreduce_densify_state(st)
    out.put((rank, imgs, float(flat[0]), st["grad2d"].tolist(), st["count"].tolist(), st["vis"].tolist(), st["radii"].tolist()))
    dist.destroy_process_group()

def _launch_world2(worker, sort_key=None, attempts=3):
    """two spawned gloo ranks on a free local port -> their queue outputs, sorted. The port is found by bind-and-release, so another
    process can take it before the ranks rendezvous: retry with a new port if a rank dies or never reports."""
    import queue as _queue
    last = None
    for _ in range(attempts):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
        for p in procs:
            p.start()
        try:
            res = [q.get(timeout=120) for _ in procs]
            for p in procs:
                p.join(timeout=60)
            return sorted(res, key=sort_key)
        except _queue.Empty as e:
            last = e
            for p in procs:
                p.kill()
                p.join(timeout=10)
    raise AssertionError(f"no result from the gloo ranks after {attempts} attempts: {last!r}")


def test_data_parallel_host_logic_world2():
    res = _launch_world2(_worker)
    (r0, im0, f0, g0, c0, v0, rad0), (r1, im1, f1, g1, c1, v1, rad1) = res
    # same epoch permutation on both ranks, disjoint images inside an epoch (5 steps x 2 ranks = the 10 images once each)
    assert sorted(im0[:5] + im1[:5]) == list(range(10)) and sorted(im0[5:10] + im1[5:10]) == list(range(10))
    assert f0 == f1 == 1.5  # mean of 1 and 2
    assert g0 == g1 == [0.0, 3.0, 6.0, 9.0, 12.0] and c0 == c1 == [2.0] * 5
    assert v0 == v1 == pytest.approx([0.5, 0.9]) and rad0 == rad1 == [1.0]


def test_single_process_is_identity():
    from gssdf_b200 import parallel
    t = torch.ones(4)
    assert parallel.allreduce_flat_grad(t) is t and (t == 1).all()
    assert sorted(parallel.image_for_rank(s, 0, 1, 7) for s in range(7)) == list(range(7))


def _xchg_worker(rank, world, port, out):
    """Drives parallel.GradientExchange through three fake steps with the hook order of GsSdfStep.step: stage [A] (touches only the
    SDF segment) -> before_render -> stages [B]-[C] -> on_sdf_grads_ready -> stage [D] (writes the splat segment) -> finish_step."""
    import torch.distributed as dist

    from gssdf_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_splat, n_sdf = 1000, 300
    flat = torch.zeros(n_splat + n_sdf)
    x = parallel.GradientExchange()
    seen = []
    for step in range(3):
        flat[n_splat:].zero_()                       # step start: only the SDF segment is cleared
        flat[n_splat:] += (rank + 1) * (step + 1)    # [A] + [C] write the SDF segment
        x.before_render()                            # previous step's splat all-reduce must be complete here ...
        if step > 0:
            seen.append(float(flat[0]))              # ... so the reduced value of the previous step is visible
        flat[:n_splat].zero_()
        x.on_sdf_grads_ready(flat[n_splat:])
        flat[:n_splat] += 10 * (rank + 1) + step     # [D] writes the splat segment
        x.finish_step(flat[:n_splat])
        seen.append(float(flat[n_splat]))            # the SDF segment is reduced when finish_step returns
    x.drain()
    seen.append(float(flat[0]))
    out.put((rank, seen))
    dist.destroy_process_group()


def test_gradient_exchange_overlap_protocol_world2():
    res = _launch_world2(_xchg_worker)
    # sums over ranks {1, 2}: sdf segment (1+2)*(step+1); splat segment 10*(1+2) + 2*step
    expect = [3.0, 30.0, 6.0, 32.0, 9.0, 34.0]
    for rank, seen in res:
        assert seen == expect, (rank, seen)


class _FakeTrainer:
    """CPU stand-in with the five members parallel.DataParallelStep needs; 'Adam' is plain SGD so the expected result is easy to state."""

    def __init__(self, rank, n_splat=64, n_sdf=24):
        self.rank, self.t0 = rank, n_splat
        self.flat_grad = torch.zeros(n_splat + n_sdf)
        self.params = torch.zeros(n_splat + n_sdf)
        self.calls = []

    def train_step(self, step, on_sdf_grads_ready=None, before_render=None):
        g = self.flat_grad
        assert float(g[self.t0:].abs().max()) == 0.0, "the SDF segment must have been consumed (zeroed) by adam_sdf"
        g[self.t0:] += (self.rank + 1) * (step + 1)                       # [A]: SDF stage on the ray samples
        if before_render is not None:
            before_render()                                               # previous splat all-reduce + splat Adam complete here
        assert float(g[:self.t0].abs().max()) == 0.0, "the splat segment must have been consumed by adam_splat"
        g[self.t0:] += 0.5 * (self.rank + 1)                              # [C]
        if on_sdf_grads_ready is not None:
            on_sdf_grads_ready(g[self.t0:])
        g[:self.t0] += 10 * (self.rank + 1) + step                        # [D]
        return step

    def _sgd(self, sl, scale):
        self.params[sl] -= 0.1 * scale * self.flat_grad[sl]
        self.flat_grad[sl] = 0

    def adam_sdf(self, scale=1.0):
        self.calls.append("sdf")
        self._sgd(slice(self.t0, None), scale)

    def adam_splat(self, scale=1.0):
        self.calls.append("splat")
        self._sgd(slice(0, self.t0), scale)

    def adam_all(self, scale=1.0):
        self.calls.append("all")
        self._sgd(slice(0, None), scale)


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist

    from gssdf_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = _FakeTrainer(rank)
    DP = parallel.DataParallelStep(T)
    for step in range(3):
        DP.step(step)
    DP.flush()
    out.put((rank, T.params.clone(), list(T.calls)))
    dist.destroy_process_group()


def test_data_parallel_step_matches_mean_gradient_world2():
    """DataParallelStep on 2 gloo ranks == one process stepping with the rank-averaged gradient; both replicas end bit-identical."""
    from gssdf_b200 import parallel
    res = _launch_world2(_dp_worker, sort_key=lambda r: r[0])
    # single-process reference with the mean of the two ranks' gradients
    ref = torch.zeros(64 + 24)
    for step in range(3):
        g = torch.zeros(64 + 24)
        for rank in range(2):
            g[64:] += ((rank + 1) * (step + 1) + 0.5 * (rank + 1)) / 2
            g[:64] += (10 * (rank + 1) + step) / 2
        ref -= 0.1 * g
    for rank, params, calls in res:
        assert torch.allclose(params, ref, rtol=1e-6, atol=1e-7), (rank, params[:2], ref[:2], params[-2:], ref[-2:])
        assert calls == ["sdf", "splat", "sdf", "splat", "sdf", "splat"], calls
    assert torch.equal(res[0][1], res[1][1])
    # single process: one fused update per step
    T = _FakeTrainer(0)
    DP = parallel.DataParallelStep(T, world=1)
    DP.step(0)
    DP.flush()
    assert T.calls == ["all"]
