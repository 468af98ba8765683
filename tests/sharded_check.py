"""One sharded optimize step on `world` ranks checked against the single-process oracle.

Runs two ways (same orchestration code, recommendation-models_b200/sharded.py):
  * CPU, gloo, NumpyOps (oracle arithmetic)      -- tests/test_sharded_cpu.py, world_size 2
  * GPU, nccl, GpuOps (libb200rec)               -- tests/test_gpu_sharded.py via torchrun
Each rank checks the rows IT owns: distinct ids, summed embedding / weight gradients, and the
allreduced dense gradients, against the oracle run on every rank's batch and summed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

KIND, FC, CIN = "deepfm", [32, 16], []
F, K, B, ROWS = 39, 16, 64, 39 * 256
SEED_DATA, SEED_PARAMS = 1234, 42


class NumpyOps:
    """The arithmetic of one rank done by the oracle on CPU tensors (test double of GpuOps)."""

    def __init__(self, torch, synth, refport, spec, cap):
        import contextlib
        self.torch, self.synth, self.refport, self.spec, self.cap = torch, synth, refport, spec, cap
        self.ctx = contextlib.nullcontext
        q = np.arange(spec.rows_local)
        gid = spec.global_id(spec.rank, q)
        self.E = synth.table_rows(SEED_PARAMS, gid, K)
        self.wt = synth.wtable_rows(SEED_PARAMS, gid)
        self.model = refport.Model(KIND, F, K, FC, CIN)
        self.mats = synth.init_mats(SEED_PARAMS, self.model.mats_size())
        self.dense = torch.zeros(self.mats.size + 1)   # [mats grad | bias grad], like the library
        self.n_unique = 0
        self.int32, self.float32 = torch.int32, torch.float32

    def empty(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype)

    def stream_ctx(self):
        return self.ctx()

    def plan(self, feats, send_ids, dst):
        f = feats.numpy().astype(np.int64)
        own = self.spec.owner(f)
        order = np.argsort(own, kind="stable")
        s = send_ids.numpy(); s[:] = -1
        d = dst.numpy()
        start = np.searchsorted(own[order], np.arange(self.spec.world))
        for p, i in enumerate(order):
            o = own[i]
            slot = p - start[o]
            assert slot < self.cap
            s[o * self.cap + slot] = f[i] // self.spec.world
            d[i] = o * self.cap + slot

    def lookup(self, recv_ids, rows, w):
        ids = recv_ids.numpy()
        ok = ids >= 0
        r = rows.numpy().reshape(-1, K); r[:] = 0
        r[ok] = self.E[ids[ok]]
        ww = w.numpy(); ww[:] = 0
        ww[ok] = self.wt[ids[ok]]

    def step_rows(self, dst, rows, w, targets, grad_rows, grad_w):
        d = dst.numpy()
        emb = rows.numpy().reshape(-1, K)[d].reshape(-1).copy()
        wn = w.numpy()[d].copy()
        bias = np.array([0.1], np.float32)
        mats = self.mats.copy()
        index = np.repeat(np.arange(B, dtype=np.int32), F)
        self.loss = self.model.backward(B, index, wn, bias, emb, mats, targets.numpy())
        grad_rows.numpy().reshape(-1, K)[d] = emb.reshape(-1, K)
        grad_w.numpy()[d] = wn
        self.dense.numpy()[:-1] = mats
        self.dense.numpy()[-1] = bias[0]

    def dense_grads(self):
        return self.dense

    def segsum_sort(self, recv_ids, unique):
        pass

    def segsum(self, recv_ids, grad_rows, grad_w, unique, G, gw):
        ids = recv_ids.numpy()
        ok = ids >= 0
        u, g = self.refport.make_embedding_grad(grad_rows.numpy().reshape(-1, K)[ok].reshape(-1), ids[ok], K)
        _, g1 = self.refport.make_weights_grad(grad_w.numpy()[ok], ids[ok])
        self.n_unique = len(u)
        unique.numpy()[:len(u)] = u
        G.numpy().reshape(-1, K)[:len(u)] = g
        gw.numpy()[:len(u)] = g1


def expected(synth, refport, spec):
    """Oracle: every rank's batch against the global table, gradients summed over ranks."""
    model = refport.Model(KIND, F, K, FC, CIN)
    mats0 = synth.init_mats(SEED_PARAMS, model.mats_size())
    tot_e, tot_w = {}, {}
    gm, gb = np.zeros_like(mats0, dtype=np.float64), 0.0
    losses = []
    for r in range(spec.world):
        index, feats = synth.make_feats(SEED_DATA, r, B, F, ROWS)
        targets = synth.make_targets(SEED_DATA, feats, B, F)
        emb = synth.table_rows(SEED_PARAMS, feats, K).reshape(-1)
        w = synth.wtable_rows(SEED_PARAMS, feats)
        bias, mats = np.array([0.1], np.float32), mats0.copy()
        losses.append(model.backward(B, index, w, bias, emb, mats, targets))
        gm += mats
        gb += bias[0]
        for i, fid in enumerate(feats):
            tot_e[fid] = tot_e.get(fid, 0) + emb[i * K:(i + 1) * K].astype(np.float64)
            tot_w[fid] = tot_w.get(fid, 0) + float(w[i])
    return tot_e, tot_w, gm, gb, losses


def run(rank, world, backend, device=None):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    from oracle import refport
    from recommendation_models_b200.sharded import GpuOps, P2PShardedParRecModel, ShardedParRecModel, ShardSpec
    synth = pkg.synth
    spec = ShardSpec(ROWS, world, rank, mode="range" if backend == "p2p_range" else "mod")
    # contiguous ranges put whole fields on one rank: buckets are unbalanced, size them for the worst case
    cap = B * F if backend == "p2p_range" else int(B * F / world * 1.5) + 64
    _, feats = synth.make_feats(SEED_DATA, rank, B, F, ROWS)
    targets = synth.make_targets(SEED_DATA, feats, B, F)
    if backend == "gloo":
        ops = NumpyOps(torch, synth, refport, spec, cap)
        tf, tt = torch.from_numpy(feats.copy()), torch.from_numpy(targets.copy())
    else:
        dev = torch.device("cuda", device)
        torch.cuda.set_device(dev)
        model = pkg.make_model(KIND, F, K, FC, CIN, device=device)
        table = pkg.EmbeddingTable(spec.rows_local, K, device=device)
        pkg._lib.check(pkg.lib().b200rec_table_init_uniform_sharded(table.handle, SEED_PARAMS, -0.05, 0.05, rank,
                                                                    world, spec.period))
        pkg.ParRecModel(model, table).setParams(np.array([0.1], np.float32),
                                                synth.init_mats(SEED_PARAMS, model.getMatsSize()))
        ops = GpuOps(pkg, model, table, spec, B, cap, torch, dev)
        tf, tt = torch.from_numpy(feats).to(dev), torch.from_numpy(targets).to(dev)
    cls = P2PShardedParRecModel if backend.startswith("p2p") else ShardedParRecModel
    if backend == "p2p_graph_twoshot":     # the >= 4 GPU form of the dense allreduce, forced on any world size
        os.environ["B200REC_TWO_SHOT"] = "1"
    else:
        os.environ.pop("B200REC_TWO_SHOT", None)
    sh = cls(ops, dist, spec, B, F, K, cap=cap)
    if backend.startswith("p2p_graph"):
        # replay mode: step 1 runs call by call (buffers take their size), steps 2 and 3 capture the
        # two parities, steps 4 and 5 replay them; the next batch is staged and sorted one step ahead
        sh.load(tf, tt)
        for _ in range(5):
            sh.load(tf, tt)
            sh.step()
        assert len(sh.graphs) >= 2
    else:
        sh.optimize(tf, tt)
    if backend == "p2p":   # a second step exercises the double-buffered id slots and the step flags
        sh.optimize(tf, tt)
    if backend != "gloo":
        torch.cuda.synchronize()
        assert int(ops.overflow.item()) == 0
        U = int(ops.n_unique.item())
        loss = float(ops.loss().item())
    else:
        U, loss = ops.n_unique, ops.loss
    uniq = sh.unique[:U].cpu().numpy()
    G = sh.G[:U * K].cpu().numpy().reshape(U, K)
    gw = sh.gw[:U].cpu().numpy()
    dense = ops.dense_grads().cpu().numpy()
    gb, gm = float(dense[-1]), dense[:-1]
    tot_e, tot_w, egm, egb, losses = expected(synth, refport, spec)
    # ids this rank owns, ascending local row
    owned = sorted(fid for fid in tot_e if spec.owner([fid])[0] == rank)
    exp_rows = np.array([int(spec.local_row([fid])[0]) for fid in owned])
    order = np.argsort(exp_rows)
    assert np.array_equal(uniq, exp_rows[order]), "distinct owned rows differ"
    assert np.array_equal(spec.global_id(rank, uniq), np.array(owned)[order])
    eG = np.array([tot_e[owned[j]] for j in order])
    eW = np.array([tot_w[owned[j]] for j in order])
    tol = lambda want: 2e-5 * np.abs(want).max() + 1e-12
    assert np.abs(G - eG).max() <= tol(eG), ("emb grad", np.abs(G - eG).max(), np.abs(eG).max())
    assert np.abs(gw - eW).max() <= tol(eW), "w grad"
    assert np.abs(gm - egm).max() <= tol(egm), ("mats grad", np.abs(gm - egm).max(), np.abs(egm).max())
    assert abs(gb - egb) <= 2e-5 * abs(egb) + 1e-6, "bias grad"
    assert abs(loss - losses[rank]) <= 1e-5 * abs(losses[rank]), "loss"
    return U


def _gpu_setup(pkg, torch, spec, rank, world, device, cap):
    from recommendation_models_b200.sharded import GpuOps
    synth = pkg.synth
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    model = pkg.make_model(KIND, F, K, FC, CIN, device=device)
    table = pkg.EmbeddingTable(spec.rows_local, K, device=device)
    pkg._lib.check(pkg.lib().b200rec_table_init_uniform_sharded(table.handle, SEED_PARAMS, -0.05, 0.05, rank,
                                                                world, spec.period))
    ps = pkg.ParRecModel(model, table)
    ps.setParams(np.array([0.1], np.float32), synth.init_mats(SEED_PARAMS, model.getMatsSize()))
    return dev, model, table, ps, GpuOps(pkg, model, table, spec, B, cap, torch, dev)


def run_adam(rank, world, device, graphed=True, steps=4):
    """`steps` sharded training steps with Adam (the reference examples' optimizer,
    rec/example/DeepFMLocalExample.scala:32) on the owner-side rows and the replicated dense params,
    replayed as CUDA graphs (the update count comes from the device step counter), against the oracle
    run on the global table: every rank's batch, gradients summed over ranks, one update per step."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    from oracle import refport
    from recommendation_models_b200.sharded import P2PShardedParRecModel, ShardSpec
    synth = pkg.synth
    spec = ShardSpec(ROWS, world, rank)
    cap = int(B * F / world * 1.5) + 64
    dev, model, table, ps, ops = _gpu_setup(pkg, torch, spec, rank, world, device, cap)
    sh = P2PShardedParRecModel(ops, dist, spec, B, F, K, cap=cap)
    sh.use_graph = graphed
    lr = 0.01
    sh.set_optimizer("adam", lr)
    batches = []
    for t in range(steps + 1):
        _, f = synth.make_feats(SEED_DATA, t * world + rank, B, F, ROWS)
        batches.append((torch.from_numpy(f).to(dev), torch.from_numpy(synth.make_targets(SEED_DATA, f, B, F)).to(dev)))
    sh.load(*batches[0])
    for t in range(steps):
        sh.load(*batches[t + 1])
        sh.step()
    sh.check()
    if graphed:
        assert len(sh.graphs) >= 2
    # ---- oracle on the global table ----------------------------------------------------------------------
    ids_all = np.arange(ROWS)
    E = synth.table_rows(SEED_PARAMS, ids_all, K)
    W = synth.wtable_rows(SEED_PARAMS, ids_all)
    o = refport.Model(KIND, F, K, FC, CIN)
    mats = synth.init_mats(SEED_PARAMS, o.mats_size())
    bias = np.array([0.1], np.float32)
    stE, stW, stM, stB = {}, {}, {}, {}
    for t in range(steps):
        GE, GW = np.zeros_like(E), np.zeros_like(W)
        gm_tot, gb_tot = np.zeros_like(mats), np.zeros_like(bias)
        touched = np.zeros(ROWS, bool)
        for r in range(world):
            index, f = synth.make_feats(SEED_DATA, t * world + r, B, F, ROWS)
            tg = synth.make_targets(SEED_DATA, f, B, F)
            emb, w = E[f].reshape(-1).copy(), W[f].copy()
            gb, gm = bias.copy(), mats.copy()
            o.backward(B, index, w, gb, emb, gm, tg)
            u, G = refport.make_embedding_grad(emb, f, K)
            _, gw = refport.make_weights_grad(w, f)
            GE[u] += G
            GW[u] += gw
            touched[u] = True
            gm_tot += gm
            gb_tot += gb
        u = np.nonzero(touched)[0]
        sub = lambda st: {k: v[u] for k, v in st.items()}
        se, sw = sub(stE), sub(stW)
        E[u] = refport.optimizer_update("adam", E[u], GE[u], se, lr, step=t + 1)
        W[u] = refport.optimizer_update("adam", W[u], GW[u], sw, lr, step=t + 1)
        for st, s_new, shape in ((stE, se, E.shape), (stW, sw, W.shape)):
            for k, v in s_new.items():
                st.setdefault(k, np.zeros(shape, np.float32))[u] = v
        mats = refport.optimizer_update("adam", mats, gm_tot, stM, lr, step=t + 1)
        bias = refport.optimizer_update("adam", bias, gb_tot, stB, lr, step=t + 1)
    gE, gW = table.read(0, spec.rows_local)
    gid = spec.global_id(rank, np.arange(spec.rows_local))
    ok = gid < ROWS
    tol = lambda want: 1e-3 * np.abs(want).max()
    assert np.abs(gE[ok] - E[gid[ok]]).max() <= tol(E), ("table", np.abs(gE[ok] - E[gid[ok]]).max())
    assert np.abs(gW[ok] - W[gid[ok]]).max() <= tol(W), "weights"
    gbias, gmats = ps.getParams()
    assert np.abs(gmats - mats).max() <= tol(mats), ("mats", np.abs(gmats - mats).max())
    assert abs(gbias[0] - bias[0]) <= 1e-3 * abs(bias[0]) + 1e-5, "bias"
    return int(ok.sum())


def run_overflow(rank, world, device):
    """A bucket capacity that cannot hold the batch's distinct ids: the step must not alias another id's
    slot, and check() must raise (on every rank, as every rank overflows)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    from recommendation_models_b200.sharded import P2PShardedParRecModel, ShardSpec
    synth = pkg.synth
    spec = ShardSpec(ROWS, world, rank)
    cap = 16                                     # 64 * 39 non-zeros -> hundreds of distinct ids per owner
    dev, model, table, ps, ops = _gpu_setup(pkg, torch, spec, rank, world, device, cap)
    sh = P2PShardedParRecModel(ops, dist, spec, B, F, K, cap=cap)
    sh.check_every = 0
    _, f = synth.make_feats(SEED_DATA, rank, B, F, ROWS)
    tf, tt = torch.from_numpy(f).to(dev), torch.from_numpy(synth.make_targets(SEED_DATA, f, B, F)).to(dev)
    sh.optimize(tf, tt)
    try:
        sh.check()
    except RuntimeError as e:
        assert "overflow" in str(e)
    else:
        raise AssertionError("bucket overflow was not reported")
    sh.check()                                   # the flag is cleared once reported
    loss = float(ops.loss().item())
    assert np.isfinite(loss)
    return cap


def _cpu_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, run(rank, world, "gloo")))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "ERR " + traceback.format_exc()))
    dist.destroy_process_group()


if __name__ == "__main__":
    # torchrun --nproc-per-node N tests/sharded_check.py   (GPU, nccl)
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for backend in ("nccl", "p2p", "p2p_graph", "p2p_graph_twoshot", "p2p_range"):
        u = run(rank, world, backend, device=local)
        print(f"rank {rank}/{world}: sharded step ok ({backend}), {u} owned distinct rows", flush=True)
    u = run_adam(rank, world, local)
    print(f"rank {rank}/{world}: sharded step ok (p2p_graph + adam, 4 steps), {u} owned rows compared", flush=True)
    u = run_overflow(rank, world, local)
    print(f"rank {rank}/{world}: sharded step ok (overflow raised at cap {u})", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)
