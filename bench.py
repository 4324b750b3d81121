#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the BRUNO hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # our arm  (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                     # reference arm: the CPU restatement of the reference
                                                           # (TF-1.7 cannot run here), timed on the box's host cores

Workload (BASELINE.json): bn2_omniglot_tp training -- conv Real NVP + recurrent t-process on synthetic Omniglot-shaped
28x28 sequences, seq_len 20, 32 sequences per GPU (weak scaling: per-GPU batch fixed), one step = forward + hand-written
backward + NCCL gradient allreduce + TF-semantics RMSProp.  Also reports forward NLL-eval throughput.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = 'bn2_omniglot_tp'
R_BLOCKS, FILTERS = 8, 64


def conv_launch_flops(per_gpu_images):
    """Algorithmic FLOPs (2*MAC) of the 3x3 64->64 convolutions per forward pass: 16 per coupling, 3 couplings at
    28x28 and 7 at 14x14."""
    per_px = 2 * 9 * FILTERS * FILTERS
    return per_gpu_images * 16 * per_px * (3 * 28 * 28 + 7 * 14 * 14)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if 'Active' in v and 'Not' not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def host_threads():
    """Host threads the CPU arm may use: every core this process is allowed on.  torchrun exports OMP_NUM_THREADS=1 to its
    workers, which silently made the round-1 reference arm 13x slower at N > 1; the arm sets its thread count itself."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_line(args, steps, warmup, as_reference_arm):
    """The reference's own CPU path cannot run (TF-1.7 absent): time the oracle restatement (PyTorch-CPU fp32, all host
    threads) on a bounded sample of the same workload -- `sample_b` sequences of `seq_len` frames per step, one step =
    forward + backward + TF-1.7 RMSProp update of all variables, i.e. the same work per sequence as our arm's step."""
    n_thr = host_threads()
    os.environ['OMP_NUM_THREADS'] = str(n_thr)
    os.environ['MKL_NUM_THREADS'] = str(n_thr)
    import torch
    torch.set_num_threads(n_thr)
    from oracle import bruno_oracle as O
    name = args.config_name if args.config_name != 'm1_shapenet' else CONFIG
    spec = O.spec_for(name)
    sample_b = args.cpu_sample_seqs
    P = O.init_params(spec, seed=0, dtype=torch.float32)
    slots = {k: torch.ones_like(v) for k, v in P.items()}
    xs = [O.synthetic_batch(spec, sample_b, args.seq_len, seed=1 + i, dtype=torch.float32) for i in range(2)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, grads = O.loss_and_grads(spec, P, xs[i % 2], student_grad_scale=0.1)
        for k in P:
            P[k], slots[k] = O.rmsprop_tf17_step(P[k], grads[k], slots[k], 1e-5)
        times.append(time.perf_counter() - t0)
    timed = sorted(times[warmup:])
    t = sum(timed) / max(steps, 1)
    val = sample_b / t
    cb = {'value': val, 'unit': 'seqs/s', 'cores': torch.get_num_threads(), 'kind': 'port',
          'median_ms_per_step': 1e3 * timed[len(timed) // 2],
          'sample': '%d sequences x %d frames per step, forward + backward + RMSProp of %s (oracle restatement of the '
                    'reference, PyTorch-CPU fp32, %d threads), %d timed steps after %d warm-up' % (
                        sample_b, args.seq_len, name, torch.get_num_threads(), steps, warmup)}
    if not as_reference_arm:
        return cb
    return {'metric': '%s_train_seqs_per_sec' % name, 'value': val, 'unit': 'seqs/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': '%s training step (fwd + bwd + RMSProp), seq_len %d, CPU restatement of the reference' % (name, args.seq_len),
                       'per_step_sample_seqs': sample_b, 'host_threads': torch.get_num_threads()},
            'cpu_baseline': cb, 'e2e': {'value': val, 'unit': 'seqs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


def dense_coupling_sequence_graph_ms(cfg, dev, rows, reps=10):
    """GPU time of the dense couplings of ONE training step: bruno_coupling_dense_fwd for every dense coupling, then
    bruno_coupling_dense_bwd in reverse order (the calls nn_extra_nvp._DenseCouplingFn makes, with N = `rows`, D = latent
    dimension, U = hidden units) as one dependent launch sequence replayed from a CUDA graph captured on a stream whose
    workspace is bound.  Returns (ms per step, kernel launches of the sequence, useful GEMM FLOPs)."""
    import torch
    from bruno_b200 import _lib
    layers = list(getattr(cfg, 'nvp_dense_layers', []))
    if not layers:
        return 0.0, 0, 0.0
    D = 1
    for v in cfg.obs_shape[1:]:
        D *= int(v)
    U = int(getattr(layers[0], 'n_units', 512))
    N = int(rows)
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(dev)
    P = dict(V1=rnd(D, U), g1=torch.ones(U, device=dev), b1=torch.zeros(U, device=dev), V2=rnd(U, U), g2=torch.ones(U, device=dev),
             b2=torch.zeros(U, device=dev), Ks=rnd(U, D), bs=torch.zeros(D, device=dev), Kt=rnd(U, D), bt=torch.zeros(D, device=dev))
    G = {k: torch.empty_like(v) for k, v in P.items()}
    x, dy, dx = rnd(N, D), rnd(N, D), torch.empty(N, D, device=dev)
    y, ld, ld2, dld = torch.empty(N, D, device=dev), torch.zeros(N, device=dev), torch.empty(N, device=dev), torch.ones(N, device=dev)
    ws = [torch.empty(_lib.query('bruno_coupling_dense_ws_floats', N, D, U), device=dev) for _ in layers]
    ws2 = torch.empty(_lib.query('bruno_coupling_dense_bwd_ws_floats', N, D, U), device=dev)
    # useful FLOPs of the 12 products of a coupling (4 forward, 8 backward)
    flops = 2.0 * N * (3 * (D * U + U * U + 2 * U * D)) * len(layers)

    def seq():
        for i, _ in enumerate(layers):
            _lib.call('bruno_coupling_dense_fwd', x, ld, y, ld2, N, D, U, 4 + (i & 1), 0, P['V1'], P['g1'], P['b1'], P['V2'], P['g2'],
                      P['b2'], P['Ks'], P['bs'], P['Kt'], P['bt'], ws[i])
        for i in reversed(range(len(layers))):
            _lib.call('bruno_coupling_dense_bwd', x, dy, dld, N, D, U, 4 + (i & 1), P['V1'], P['g1'], P['b1'], P['V2'], P['g2'], P['b2'],
                      P['Ks'], P['Kt'], ws[i], ws2, dx, G['V1'], G['g1'], G['b1'], G['V2'], G['g2'], G['b2'], G['Ks'], G['bs'], G['Kt'],
                      G['bt'])

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        seq()
        torch.cuda.synchronize()
        l0 = _lib.query('bruno_launch_count')
        seq()
        launches = _lib.query('bruno_launch_count') - l0
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        seq()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, launches, flops


def conv_sequence_graph_ms(cfg, dev, reps=10, direction=None, height=None):
    """GPU time of the 3x3 conv launches of ONE training step in the regime the timed region runs them in: the same
    kernels, modes, shapes and data dependencies (per coupling: 2R forward launches as a dependent chain, then 2R dgrad
    launches with their saved-activation / addend tiles), captured into one CUDA graph and replayed; CUDA events on the
    launch stream around the replays.  (Per-launch event spans need eager launches, whose 14x14 layers are host-launch
    bound and whose events break the programmatic-dependent-launch overlap: they overstate a layer by ~30 %.)"""
    import torch
    from bruno_b200 import _lib, slab
    plan = []
    for layer in cfg.nvp_layers:
        sl = getattr(layer, '_slabs', None)
        if sl is None:
            continue
        (N, H, W, _), R = sl[0], layer.num_res_blocks
        wb = _lib.query('bruno_conv_wpack_bytes')
        wf = torch.empty(2 * R, wb, dtype=torch.uint8, device=dev)
        wbw = torch.empty(2 * R, wb, dtype=torch.uint8, device=dev)
        _lib.call('bruno_conv_wn_pack', layer.Vr.detach(), layer.gr.detach(), wf, wbw, 2 * R)
        if height is not None and H != height:                   # tools/bench_conv_step.py: one resolution at a time
            continue
        plan.append((N, H, W, R, sl[1], wf, wbw, layer.br.detach(), slab.alloc_slabs(3, N, H, W, dev)))

    def run():
        n = 0
        for N, H, W, R, saved, wf, wbw, br, tmp in (plan if direction != 'bwd' else []):   # forward: h -> c2 -> c3 (+ skip) -> h'
            hi = 0
            for r in range(R):
                ui, ni = (hi + 1) % 3, (hi + 2) % 3
                _lib.call('bruno_conv3x3', 0, tmp[hi], wf[2 * r], br[2 * r].contiguous(), 0, None, None, tmp[ui], N, H, W)
                _lib.call('bruno_conv3x3', 1, tmp[ui], wf[2 * r + 1], br[2 * r + 1].contiguous(), 0, tmp[hi], None, tmp[ni], N, H, W)
                hi = ni
                n += 2
        for N, H, W, R, saved, wf, wbw, br, tmp in (reversed(plan) if direction != 'fwd' else []):  # backward: dgrad chain per block
            for r in range(R - 1, -1, -1):
                # g2 = dgrad_c3(g3) * ELU'(u_r);  g3' = (dgrad_c2(g2) + g3) * ELU'(h_r): saved activations as aux tiles
                g3, g2, g1 = (0, 1, 2) if r % 2 == 0 else (2, 1, 0)
                _lib.call('bruno_conv3x3', 2, tmp[g3], wbw[2 * r + 1], None, 0, saved[2 * r + 1], None, tmp[g2], N, H, W)
                _lib.call('bruno_conv3x3', 3, tmp[g2], wbw[2 * r], None, 0, saved[2 * r], tmp[g3], tmp[g1], N, H, W)
                n += 2
        return n
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        launches = run()
    for _ in range(2):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, launches


def bench_conditional(args, dev, world, rank):
    """Conditional BRUNO (config_conditional/m1_shapenet: B=4, T=16, 32x32x1, label dim 2, GP with n_context=1), BASELINE.json
    config C5: training step + NLL evaluation through the same public calls a user of the config module makes."""
    import torch
    import torch.distributed as dist
    from bruno_b200 import _lib
    from bruno_b200.config_conditional import defaults
    defaults.seq_len, defaults.n_context = 16, 1
    cfg = importlib.import_module('bruno_b200.config_conditional.m1_shapenet')
    from bruno_b200.config_conditional.train import ConditionalTrainer
    it = cfg.train_data_iter.generate()
    batches = [next(it) for _ in range(4)]
    hx = [torch.from_numpy(b[0]).pin_memory() for b in batches]
    hy = [torch.from_numpy(b[1]).pin_memory() for b in batches]
    xs, ys = [h.to(dev) for h in hx], [h.to(dev) for h in hy]
    tr = ConditionalTrainer(cfg, dev, world_size=world, rank=rank).materialize(xs[0], ys[0], init=True)
    warmup, steps = max(args.warmup, 3), args.steps

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    losses = []

    def e2e(i):
        x, y = hx[i % 4].to(dev, non_blocking=True), hy[i % 4].to(dev, non_blocking=True)
        losses.append(tr.train_step(x, y, 1e-3, 0.5).item())
    for i in range(warmup):
        tr.train_step(xs[i % 4], ys[i % 4], 1e-3, 0.5)
    l0 = _lib.query('bruno_launch_count')
    ms_train = timed(lambda i: tr.train_step(xs[i % 4], ys[i % 4], 1e-3, 0.5), steps)
    launches = _lib.query('bruno_launch_count') - l0
    ms_e2e = timed(e2e, steps)
    ms_eval = timed(lambda i: tr.eval_step(xs[i % 4], ys[i % 4]), steps)
    B = cfg.batch_size * world
    return {'metric': 'm1_shapenet_train_seqs_per_sec', 'value': B * steps / (ms_train * 1e-3), 'unit': 'seqs/s', 'n_gpus': world,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms_train / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'Conditional BRUNO m1_shapenet training step: 14 label-conditioned conv couplings (R=6, 3x3 c1) + 6 '
                                   'label-conditioned dense couplings, GP, n_context 1; fwd + bwd + RMSProp',
                       'obs_shape': list(cfg.obs_shape[1:]), 'seq_len': cfg.seq_len, 'seqs_per_gpu': cfg.batch_size, 'params': tr.n_params},
            'eval_nll': {'value': B * steps / (ms_eval * 1e-3), 'unit': 'seqs/s', 'ms_per_step': ms_eval / steps},
            'e2e': {'value': B * steps / (ms_e2e * 1e-3), 'unit': 'seqs/s', 'ms_per_step': ms_e2e / steps,
                    'h2d_bytes_per_step': int(hx[0].numel() * 4 + hy[0].numel() * 4), 'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches), 'final_loss': losses[-1] if losses else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='bruno_b200')
    ap.add_argument('--batch', type=int, default=32, help='sequences per GPU')
    ap.add_argument('--seq_len', type=int, default=20)
    ap.add_argument('--cpu_sample_seqs', type=int, default=2, help='sequences per CPU step (BASELINE.md section 5: B = 2)')
    ap.add_argument('--no_cpu_baseline', action='store_true')
    ap.add_argument('--config_name', default=CONFIG, help='config_rnn module; the headline metric is quoted on the default')
    ap.add_argument('--no_graph', action='store_true', help='time the training step with eager launches instead of the CUDA graph')
    ap.add_argument('--strong', action='store_true', help='strong scaling: --batch is the GLOBAL batch, split over the ranks (train.py:184)')
    ap.add_argument('--data_kind', default=None, help="synthetic pixel distribution: 'strokes' (Omniglot-like) or 'bytes' (uniform)")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    # exactly ONE line may reach stdout: route everything else (NCCL banner, config prints) to stderr
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    def emit(obj):
        real_stdout.write(json.dumps(obj) + '\n')
        real_stdout.flush()

    if args.impl == 'reference':
        if rank == 0:
            emit(cpu_reference_line(args, args.steps, args.warmup, True))
        return

    import torch
    import torch.distributed as dist
    warmup = max(args.warmup, 3)
    steps = args.steps
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    from bruno_b200 import _lib
    if args.config_name == 'm1_shapenet':
        line = bench_conditional(args, dev, world, rank)
        if rank == 0:
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return
    from bruno_b200.config_rnn import defaults
    from bruno_b200.trainer import Trainer
    defaults.seq_len = args.seq_len
    cfg = importlib.import_module('bruno_b200.config_rnn.' + args.config_name)
    if args.strong:
        assert args.batch % world == 0, '--strong: the global batch must be divisible by the number of GPUs'
        per_gpu = args.batch // world
    else:
        per_gpu = args.batch
    cfg.batch_size = per_gpu * world                         # reference semantics: global batch, split over towers
    B, T = per_gpu, args.seq_len
    H, W, C = cfg.obs_shape[1:]

    from bruno_b200.synthetic_data import SyntheticExchSeqDataIterator
    import numpy as np
    kind = args.data_kind or ('bytes' if 'cifar' in args.config_name else 'strokes')
    it = SyntheticExchSeqDataIterator(seq_len=T, batch_size=B, obs_shape=(H, W, C), kind=kind,
                                      rng=np.random.RandomState(42 + rank), noise_rng=np.random.RandomState(317070 + rank))
    gen = it.generate()
    n_host = 4
    host = [torch.from_numpy(next(gen)).pin_memory() for _ in range(n_host)]
    devx = [h.to(dev) for h in host]

    tr = Trainer(cfg, dev, world_size=world, rank=rank).materialize(devx[0], init=True)
    # the reference's heads are zero-initialised (flow = identity at step 0); a few real steps move them off zero
    sched = getattr(cfg, 'learning_rate_schedule', None)
    lr = sched[200] if sched and 200 in sched else cfg.learning_rate

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    losses = []

    def train_resident(i):
        losses.append(tr.train_step(devx[i % n_host], lr, 0.1))

    def train_e2e(i):
        x = host[i % n_host].to(dev, non_blocking=True)
        loss = tr.train_step(x, lr, 0.1)
        losses.append(loss.item())                          # device->host read of the step's result

    def eval_resident(i):
        tr.eval_step(devx[i % n_host])

    for i in range(warmup):
        train_resident(i)
    torch.cuda.synchronize()
    l0 = _lib.query('bruno_launch_count')
    train_resident(0)                                       # one eager step: how many of our kernels a step launches
    launches_per_step = _lib.query('bruno_launch_count') - l0
    if not args.no_graph:
        # forward + hand-written backward as ONE CUDA-graph launch (Trainer.enable_cuda_graph); allreduce + optimiser eager
        tr.enable_cuda_graph(devx[0])
        for i in range(2):
            train_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    torch.cuda.cudart().cudaProfilerStart()                 # `ncu --profile-from-start off` captures exactly the timed region
    ms_train = timed(train_resident, steps)
    torch.cuda.cudart().cudaProfilerStop()
    launches = launches_per_step * steps
    ms_e2e = timed(train_e2e, steps)
    if sampler:
        sampler.stop_flag = True
    for i in range(2):
        eval_resident(i)
    ms_eval = timed(eval_resident, steps)

    # 20-way 1-shot scoring episodes (test_few_shot_omniglot: batch 20, seq_len 2, forward only) -- latency-bound case
    from bruno_b200.synthetic_data import SyntheticFewShotIterator
    fs_it = SyntheticFewShotIterator(seq_len=2, batch_size=20, obs_shape=(H, W, C), rng=np.random.RandomState(7 + rank), n_classes=24)
    fs_x = [torch.from_numpy(e[0]).to(dev) for e, _ in zip(fs_it.generate(), range(4))]

    def few_shot(i):
        with torch.no_grad():
            cfg.build_model(fs_x[i % 4], want_prior=False)[0][:, -1].argmax()

    for i in range(3):
        few_shot(i)
    fs_steps = max(steps, 10)
    ms_fs = timed(few_shot, fs_steps)
    graphed = cfg.model.graphed_forward(fs_x[0])            # the same episode as ONE CUDA-graph launch

    def few_shot_graph(i):
        graphed(fs_x[i % 4])[0][:, -1].argmax()

    for i in range(3):
        few_shot_graph(i)
    ms_fsg = timed(few_shot_graph, fs_steps)

    # second pass with per-launch CUDA events on the conv kernels (live durations inside the step)
    tr.disable_cuda_graph()                                  # per-launch CUDA events need eager launches
    _lib.query('bruno_profile_enable', 1)
    ms_train_eager = timed(train_resident, steps)
    import ctypes
    tot, cnt = ctypes.c_double(), ctypes.c_long()
    prof = {}
    for kind, name in ((0, 'conv3x3_fwd_dgrad'), (1, 'conv3x3_wgrad'), (2, 'process_fwd'), (3, 'process_bwd')):
        _lib.load().bruno_profile_read(kind, ctypes.byref(tot), ctypes.byref(cnt))
        prof[name] = {'total_ms': tot.value, 'launches': cnt.value}
    _lib.query('bruno_profile_enable', 0)
    conv_graph_ms, conv_graph_launches = conv_sequence_graph_ms(cfg, dev)
    dense_ms, dense_launches, dense_flops = dense_coupling_sequence_graph_ms(cfg, dev, B * T)

    # The reference's own nr_gpu semantics (config_rnn/train.py:184: ONE batch of `batch_size` sequences, np.split over the
    # GPUs) = strong scaling: measured in the same run with the global batch of --batch sequences split over the ranks.
    strong = None
    if world > 1 and not args.strong and args.batch % world == 0:
        per = args.batch // world
        xs_small = [d[:per].contiguous() for d in devx]

        def train_small(i):
            tr.train_step(xs_small[i % n_host], lr, 0.1)
        for i in range(3):
            train_small(i)
        if not args.no_graph:
            tr.enable_cuda_graph(xs_small[0])
            for i in range(2):
                train_small(i)
        ms_small = timed(train_small, steps)
        tr.disable_cuda_graph()
        strong = {'global_batch': args.batch, 'seqs_per_gpu': per, 'value': args.batch * steps / (ms_small * 1e-3), 'unit': 'seqs/s',
                  'ms_per_step': ms_small / steps,
                  'what': 'ONE batch of %d sequences split over %d GPUs (np.split, config_rnn/train.py:184) -- the reference\'s '
                          'nr_gpu semantics; speed-up over 1 GPU = this value / the 1-GPU bench value' % (args.batch, world)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    # Two measured denominators (MEASURED_PEAKS.json): the burst figure (a kernel timed alone / a short region at full
    # clocks) and the sustained one (inside a long, power-limited step).  Both fractions are reported; `peak` / `frac` use
    # the one that matches the SM clock sampled DURING the timed region (>= 90 % of max -> burst regime).
    peak_burst = peaks.get('bf16_tflops', 1660.8)
    peak_sust = peaks.get('bf16_tflops_sustained', 1396.3)
    clk = sampler.summary() if sampler else {}
    burst_regime = bool(clk.get('sm_mhz') and clk.get('sm_max_mhz') and clk['sm_mhz'] >= 0.9 * clk['sm_max_mhz'])
    peak_tf = peak_burst if burst_regime else peak_sust
    peak_src = ('measured (MEASURED_PEAKS.json %s; SM clock under load %s of %s MHz -> %s regime)' % (
        'bf16_tflops' if burst_regime else 'bf16_tflops_sustained', clk.get('sm_mhz'), clk.get('sm_max_mhz'),
        'burst' if burst_regime else 'sustained')) if peaks else 'fallback (B200_PROFILING.md)'
    n_img = B * T
    conv = prof['conv3x3_fwd_dgrad']
    # forward convs + dgrad convs: 2 x the forward count per step, `steps` steps in the profiled region
    if args.config_name == CONFIG:
        fwd_flops = conv_launch_flops(n_img)
    else:                                                    # 2R convolutions per conv coupling at the coupling's resolution
        fwd_flops = 0.0
        for layer in cfg.nvp_layers:
            sl = getattr(layer, '_slabs', None)
            if sl is not None:
                (n_, h_, w_, _), R_ = sl[0], layer.num_res_blocks
                fwd_flops += 2.0 * R_ * 2 * 9 * FILTERS * FILTERS * n_ * h_ * w_
    flops_total = 2.0 * fwd_flops * steps
    achieved_spans = flops_total / (conv['total_ms'] * 1e-3) / 1e12 if conv['total_ms'] > 0 else None
    achieved = 2.0 * fwd_flops / (conv_graph_ms * 1e-3) / 1e12
    traffic, traffic_note = None, None
    tj = os.path.join(ROOT, 'profiles', 'r02_conv_traffic.json')
    if os.path.exists(tj):
        tinfo = json.load(open(tj))
        traffic = tinfo['avg_bytes_per_launch']
        traffic_note = ('dram read+write bytes per 640x28x28 launch, mean of modes 0 / 1 / 3, ncu --set full (%s); measured %s; '
                        'algorithmic %s' % (tinfo['source'], tinfo['dram_bytes_per_launch'], tinfo['algorithmic_bytes_per_launch']))
    # t-process forward scan against the HBM roofline: the training shape (B = 32) is launch-bound, so the scan is also
    # timed alone on a batch far larger than L2 (16384 sequences = 1.03 GB of latents), CUDA events on the launch stream
    scan = None
    try:
        from bruno_b200.process import build_table
        from bruno_b200.nn_extra_student import StudentRecurrentLayer
        from bruno_b200.nn_extra_gauss import GaussianRecurrentLayer
        Dz, Bz = H * W * C, 16384
        lay = GaussianRecurrentLayer((Dz,)) if args.config_name.endswith('_gp') else StudentRecurrentLayer((Dz,))
        zz = torch.randn(Bz, T, Dz, device=dev)
        nu = lay.nu_vbl.detach() if lay.nu_vbl is not None else None
        tab = build_table(lay.kind, lay.var_vbl.detach(), nu, lay.corr_vbl.detach(), lay.mu, None, lay.min_nu, 0, T, False)
        lpz = torch.empty(Bz, T, device=dev)
        scr = torch.empty(max(_lib.query('bruno_process_scratch_floats', Bz, T, Dz), 1), device=dev)

        lppz = torch.empty(Bz, T, device=dev)
        tab_g = build_table(lay.kind, lay.var_vbl.detach(), nu, lay.corr_vbl.detach(), lay.mu, None, lay.min_nu, 0, T, True)
        dlp = torch.randn(Bz, T, device=dev)
        dz = torch.empty_like(zz)
        dvar, dnu, dcorr = torch.empty(1, Dz, device=dev), (torch.empty(1, Dz, device=dev) if nu is not None else None), torch.empty(1, Dz, device=dev)
        scr_b = torch.empty(max(_lib.query('bruno_process_bwd_scratch_floats', Bz, T, Dz), 1), device=dev)

        def time_it(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3 / 10

        hbm = peaks.get('hbm_gbs', 6550.0)
        kname = 'GP' if lay.nu_vbl is None else 'TP'
        # algorithmic bytes (DESIGN 5.3 / SURVEY 8d3): fwd = z once + log-probs; fwd+prior adds the prior log-probs;
        # bwd = z read + dz written + dlp read
        t_f = time_it(lambda: _lib.call('bruno_process_fwd', lay.kind, zz, lay.mu, tab, None, None, lpz, None, scr, Bz, T, Dz))
        t_fp = time_it(lambda: _lib.call('bruno_process_fwd', lay.kind, zz, lay.mu, tab, None, None, lpz, lppz, scr, Bz, T, Dz))
        t_b = time_it(lambda: _lib.call('bruno_process_bwd', lay.kind, zz, lay.mu, tab_g, dlp, None, dz, dvar, dnu, dcorr, scr_b, Bz, T, Dz))
        alg = 4.0 * Bz * T * Dz + 4.0 * Bz * T
        alg_fp = 4.0 * Bz * T * Dz + 8.0 * Bz * T
        alg_b = 8.0 * Bz * T * Dz + 4.0 * Bz * T
        t_scan = t_f
        traffic_scan = None                                   # measured DRAM bytes per latent element (ncu --set full) x elements
        tpj = os.path.join(ROOT, 'profiles', 'r01_process_traffic.json')
        if os.path.exists(tpj) and lay.nu_vbl is not None:
            traffic_scan = json.load(open(tpj))['bytes_per_latent_element'] * Bz * T * Dz
        scan = {'bound': 'hbm', 'kernel': 'process_fwd_ring_kernel (%s forward scan, persistent bulk-copy ring)' % kname,
                'achieved': alg / t_scan * 1e-9, 'peak': hbm, 'unit': 'GB/s', 'frac': alg / t_scan * 1e-9 / hbm,
                'traffic': traffic_scan, 'us_per_launch': t_scan * 1e6, 'shape': [Bz, T, Dz],
                'fwd_prior': {'what': 'forward scan with the prior log-probs (the NLL-eval path, want_prior=True)',
                              'achieved': alg_fp / t_fp * 1e-9, 'frac': alg_fp / t_fp * 1e-9 / hbm, 'us_per_launch': t_fp * 1e6},
                'bwd': {'what': 'hand-written backward scan (dz + parameter gradients)',
                        'achieved': alg_b / t_b * 1e-9, 'frac': alg_b / t_b * 1e-9 / hbm, 'us_per_launch': t_b * 1e6},
                'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback'}
        del dz
        del zz
    except Exception as exc:                                   # the bench line must not die on the side measurement
        scan = {'error': repr(exc)}

    seqs = B * world
    line = {
        'metric': '%s_train_seqs_per_sec' % args.config_name, 'value': seqs * steps / (ms_train * 1e-3), 'unit': 'seqs/s',
        'n_gpus': world, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms_train / steps, 'higher_is_better': True,
        'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': '%s training step: Real NVP (%d conv + %d dense couplings) + recurrent '
                               'process, fwd + hand-written bwd + NCCL allreduce + optimiser (TF-1.7 semantics)' % (
                                   args.config_name, sum(1 for l in cfg.nvp_layers if hasattr(l, 'num_res_blocks')), len(cfg.nvp_dense_layers)),
                   'obs_shape': [H, W, C], 'seq_len': T, 'seqs_per_gpu': B, 'global_batch': seqs, 'parallelism': 'dp%d' % world,
                   'params': tr.n_params,
                   'l2_note': 'working set per step (%.1f GB of saved bf16 activations) >> 126 MB L2; inputs rotate over %d batches' % (
                       sum(l._slabs[1].numel() for l in cfg.nvp_layers if getattr(l, '_slabs', None) is not None) / 1e9, n_host)},
        'eval_nll': {'value': seqs * steps / (ms_eval * 1e-3), 'unit': 'seqs/s', 'ms_per_step': ms_eval / steps,
                     'what': 'forward NLL (log_probs, latent, prior) of the same batch shape'},
        'few_shot_20way_1shot': {'value': world * fs_steps / (ms_fsg * 1e-3), 'unit': 'episodes/s', 'ms_per_episode': ms_fsg / fs_steps,
                                 'ms_per_episode_eager': ms_fs / fs_steps,
                                 'what': 'test_few_shot_omniglot scoring: 20 candidate sequences x 2 frames, forward + argmax'},
        'e2e': {'value': seqs * steps / (ms_e2e * 1e-3), 'unit': 'seqs/s', 'h2d_bytes_per_step': int(host[0].numel() * 4),
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / steps},
        'gpu_launches': int(launches),
        'gpu_launches_note': 'kernels of libbruno_sm100.so per step (%d, counted on an eager step) x steps; in the timed region '
                             'the forward + backward kernels are replayed from ONE CUDA graph (%s)' % (
                                 launches_per_step, 'off: --no_graph' if args.no_graph else 'on'),
        'strong_scaling': strong,
        'ms_per_step_eager': ms_train_eager / steps,
        'roofline': {'bound': 'tensor', 'kernel': 'conv3x3_line_kernel (line-tile tcgen05 implicit GEMM, fwd + dgrad launches)',
                     'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                     'frac': (achieved / peak_tf) if achieved else None,
                     'frac_burst': (achieved / peak_burst) if achieved else None,
                     'frac_sustained': (achieved / peak_sust) if achieved else None,
                     # the burst peak scaled to the SM clock sampled during the timed region (tensor throughput is per clock)
                     'frac_at_sampled_clock': (achieved / (peak_burst * clk['sm_mhz'] / clk['sm_max_mhz']))
                     if (achieved and clk.get('sm_mhz') and clk.get('sm_max_mhz')) else None,
                     'traffic': traffic, 'traffic_note': traffic_note,
                     'peak_source': peak_src,
                     'how': 'the %d fwd + dgrad conv launches of one training step (same kernels, modes, shapes, dependencies) replayed '
                            'from one CUDA graph like the timed region, CUDA events on the launch stream' % conv_graph_launches,
                     'launches': conv_graph_launches, 'avg_launch_us': 1e3 * conv_graph_ms / max(conv_graph_launches, 1),
                     'ms_per_step': conv_graph_ms, 'share_of_step': conv_graph_ms / (ms_train / steps),
                     'eager_event_spans': {'achieved': achieved_spans, 'frac_burst': (achieved_spans / peak_burst) if achieved_spans else None,
                                           'avg_launch_us': 1e3 * conv['total_ms'] / max(conv['launches'], 1),
                                           'what': 'per-launch CUDA-event spans of an EAGER step (round-1 method): the events break the '
                                                   'programmatic-dependent-launch overlap and the 14x14 layers are host-launch bound'}},
        'roofline_wgrad': (lambda w: {'bound': 'tensor', 'kernel': 'conv3x3_wgrad_line_kernel (line-tile tcgen05 wgrad, one launch per coupling)',
                                        'achieved': fwd_flops * steps / (w['total_ms'] * 1e-3) / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s',
                                        'frac': fwd_flops * steps / (w['total_ms'] * 1e-3) / 1e12 / peak_tf,
                                        'frac_burst': fwd_flops * steps / (w['total_ms'] * 1e-3) / 1e12 / peak_burst,
                                        'launches': w['launches'], 'ms_per_step': w['total_ms'] / steps,
                                        'how': 'per-launch CUDA-event spans of the eager step (a wgrad launch is 130-480 us: the events cost nothing); '
                                               'algorithmic FLOPs = those of the forward convolutions',
                                        'ncu': 'profiles/r02_wgrad_line_ncu_summary.txt: tensor pipe active 86.5 %'}
                           if w['total_ms'] > 0 else None)(prof['conv3x3_wgrad']),
        'roofline_dense': ({'bound': 'tensor',
                            'kernel': 'tf32x3_gemm_kernel (3xTF32 tcgen05 GEMM, split-K inside a thread-block cluster, paired products) '
                                      'inside bruno_coupling_dense_fwd / _bwd',
                            'achieved': dense_flops / (dense_ms * 1e-3) / 1e12, 'peak': peak_burst / 6.0, 'unit': 'TFLOP/s',
                            'frac': dense_flops / (dense_ms * 1e-3) / 1e12 / (peak_burst / 6.0),
                            'peak_source': 'derived: measured bf16 burst peak / 2 (TF32 rate) / 3 (tensor products per fp32-accurate product); '
                                           'no TF32 peak is measured on this pool',
                            'launches': dense_launches, 'ms_per_step': dense_ms, 'share_of_step': dense_ms / (ms_train / steps),
                            'how': 'ALL kernels of the dense couplings of one training step (forward calls, then backward calls in '
                                   'reverse order: 9 GEMM launches + 14 elementwise / reduction launches per coupling) replayed from one '
                                   'CUDA graph; useful FLOPs = the 12 matrix products per coupling (2MNK); the products are '
                                   'latency-bound (0.4-0.6 GFLOP each), see DESIGN 5.2b'}
                           if dense_launches else None),
        'roofline_process_scan': scan,
        'kernel_time_ms_per_step': {k: v['total_ms'] / steps for k, v in prof.items()},
        'clocks': sampler.summary() if sampler else None,
        'final_loss': float(losses[-1]) if losses else None,
    }
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_reference_line(args, 8, 2, False)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
