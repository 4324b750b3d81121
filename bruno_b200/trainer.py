"""The data-parallel training step of config_rnn/train.py:57-131 on the B200 path.

Reference semantics kept: per-tower loss = mean over the tower's shard (train.py:88-90), gradients summed over towers
and divided by nr_gpu (:96-105), gradients of the prior_* variables multiplied by the student-grad schedule value
(:108-111), RMSProp/Adam with TF-1.7 semantics (:121-131), iteration 0 = data-dependent init without a weight update
(:178-182).  B200 design: one process per GPU; the towers' sum is ONE NCCL allreduce over the flat gradient buffer
(NVLink 5 / NVSwitch); the 1/nr_gpu factor is folded into the optimiser kernel.  Optionally (`overlap_allreduce`) the
allreduce is issued per coupling layer on a side stream as soon as that layer's backward kernels are enqueued (the
gradients are written in place into the layer's contiguous slice of the flat buffer), the uncovered slices go last."""
import torch
import torch.distributed as dist

from . import _lib, nn_extra_nvp


def uncovered_ranges(done, total):
    """[(offset, length)] of [0, total) not covered by the (offset, length) pairs in `done`."""
    out, pos = [], 0
    for off, n in sorted(done) + [(total, 0)]:
        if off > pos:
            out.append((pos, off - pos))
        pos = max(pos, off + n)
    return out


class Trainer(object):
    def __init__(self, config, device, world_size=1, rank=0):
        self.cfg = config
        self.device = torch.device(device)
        self.world_size = world_size
        self.rank = rank
        self.model = None
        self.iteration = 0
        self.slots = None

    # -- variables -----------------------------------------------------------------------------------------
    def materialize(self, x_init, init=True):
        """Creates the variables from the first batch, runs the data-dependent init pass (rank 0's batch, like the
        reference which uses the full init batch on one device, train.py:60-61) and broadcasts the result."""
        cfg = self.cfg
        with torch.no_grad():
            cfg.build_model(x_init[:1], want_prior=False)           # shapes the variables
            self.model = cfg.model
            if init:
                cfg.build_model(x_init, init=True, want_prior=False)
        self.flat, self.flat_grad = self.model.flatten_parameters()
        if self.world_size > 1:
            dist.broadcast(self.flat, 0)
        opt = getattr(cfg, 'optimizer', 'adam')                        # train.py:122-131: RMSProp only if asked for
        if opt == 'rmsprop':
            self.slots = (torch.ones_like(self.flat),)                  # TF initialises the rms slot to 1
        else:
            self.slots = (torch.zeros_like(self.flat), torch.zeros_like(self.flat))
        self.opt_steps = 0
        self._layer_ranges = self._contiguous_layer_ranges()
        self._comm_stream = torch.cuda.Stream(device=self.device) if self.world_size > 1 else None
        # measured on 8 B200 (bn2_omniglot_tp, 59 MB of gradients): 14.99 ms per step with the per-layer overlapped calls vs
        # 14.86 ms with ONE allreduce after the backward -- NVSwitch moves the buffer in ~0.15 ms, there is nothing to hide,
        # and 17 small collectives cost more than one.  Off by default; kept (and tested, tools/dbg_overlap.py) for
        # configurations where the gradient buffer is large against the step.
        self.overlap_allreduce = False
        return self

    def _contiguous_layer_ranges(self):
        """id(layer) -> (offset, length) of the layer's variables in the flat buffers (flatten_parameters lays the
        variables out layer by layer; a layer whose slice were not contiguous is left to the final allreduce)."""
        index = self.model._index
        out = {}
        for layer in self.model.layers():
            names = [n for n, _ in layer.named_parameters()] if hasattr(layer, 'named_parameters') else []
            if not names:
                continue
            spans = sorted((index[n][0], index[n][0] + (index[n][1] + 3) // 4 * 4) for n in names)
            if all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1)):
                out[id(layer)] = (spans[0][0], spans[-1][1] - spans[0][0])
        return out

    @property
    def n_params(self):
        return sum(p.numel() for p in self.model.parameters())

    # -- one step ------------------------------------------------------------------------------------------
    def enable_cuda_graph(self, x_example):
        """Captures forward + hand-written backward of ONE batch shape (everything between the input copy and the gradient
        allreduce: ~600 kernel launches for bn2_omniglot_tp) into a CUDA graph; train_step then replays it with one launch.
        The allreduce and the optimiser stay outside (learning rate and student-grad scale change every iteration,
        train.py:165-170, and are kernel arguments).  What it buys: at 32 sequences per GPU the step is GPU-bound either way;
        in the reference's own nr_gpu semantics (ONE batch of 32 split over 8 GPUs = 80 images per GPU, train.py:184) the
        host cannot issue 600 launches in the ~3 ms the kernels take -- the graph removes that bound."""
        if self.overlap_allreduce:
            raise RuntimeError('the per-layer overlapped allreduce issues NCCL calls from inside the backward: not capturable')
        self._graph = None
        static_x = x_example.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                   # warm-up: slab pools, one-time kernel attributes, autograd buffers
                self._forward_backward_eager(static_x)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):          # the warmed-up stream: its GEMM workspace is bound
            static_loss = self._forward_backward_eager(static_x)
        self._graph, self._graph_x, self._graph_loss = graph, static_x, static_loss
        return self

    def disable_cuda_graph(self):
        self._graph = None

    def forward_backward(self, x):
        g = getattr(self, '_graph', None)
        if g is not None and tuple(x.shape) == tuple(self._graph_x.shape):
            self._graph_x.copy_(x, non_blocking=True)
            g.replay()
            self._reduced = []
            return self._graph_loss                   # static buffer: overwritten by the next replay
        return self._forward_backward_eager(x)

    def _forward_backward_eager(self, x):
        self.flat_grad.zero_()
        log_probs = self.cfg.build_model(x, want_prior=False)[0]
        loss = self.cfg.loss(log_probs)
        self._reduced = []                                               # [(offset, length)] already handed to NCCL
        overlap = self.world_size > 1 and self.overlap_allreduce

        def layer_done(layer):
            rng = self._layer_ranges.get(id(layer))
            if rng is None:
                return
            ready = torch.cuda.Event()
            ready.record()                                               # after the layer's backward kernels
            self._comm_stream.wait_event(ready)
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(self.flat_grad[rng[0]:rng[0] + rng[1]], op=dist.ReduceOp.SUM)
            self._reduced.append(rng)

        # the coupling backward kernels write straight into the (just zeroed) flat gradient views
        nn_extra_nvp.direct_grads(True)
        nn_extra_nvp.on_layer_grads_ready(layer_done if overlap else None)
        try:
            loss.backward()
        finally:
            nn_extra_nvp.direct_grads(False)
            nn_extra_nvp.on_layer_grads_ready(None)
        return loss.detach()

    def _allreduce_rest(self):
        """grads[0][j] += grads[i][j] (train.py:97-102) for everything the per-layer calls have not covered."""
        done = sorted(getattr(self, '_reduced', []))
        if not done:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            return
        torch.cuda.current_stream().wait_stream(self._comm_stream)
        for off, n in uncovered_ranges(done, self.flat_grad.numel()):
            dist.all_reduce(self.flat_grad[off:off + n], op=dist.ReduceOp.SUM)
        self._reduced = []

    def apply_gradients(self, lr, student_scale=1.0):
        if self.world_size > 1:
            self._allreduce_rest()
        if student_scale != 1.0:
            for off, n in self.model.student_param_ranges():             # train.py:108-111
                self.flat_grad[off:off + n].mul_(student_scale)
        gs = 1.0 / self.world_size                                       # train.py:103-105
        n = self.flat.numel()
        self.opt_steps += 1
        if len(self.slots) == 1:
            _lib.call('bruno_rmsprop_tf17', self.flat, self.flat_grad, self.slots[0], n, float(lr), 0.9, 1e-10, gs)
        else:
            _lib.call('bruno_adam_tf17', self.flat, self.flat_grad, self.slots[0], self.slots[1], n, float(lr), 0.95, 0.9995,
                      1e-8, self.opt_steps, gs)

    def train_step(self, x, lr, student_scale=1.0):
        loss = self.forward_backward(x)
        self.apply_gradients(lr, student_scale)
        return loss

    def eval_step(self, x):
        with torch.no_grad():
            return self.cfg.build_model(x, want_prior=True)

    # -- checkpoint (train.py:225-239 keeps params + optimiser slots + meta) ---------------------------------
    def state_dict(self):
        return {'variables': self.model.export_tf_variables(), 'flat': self.flat.detach().cpu(),
                'slots': [s.cpu() for s in self.slots], 'opt_steps': self.opt_steps, 'iteration': self.iteration}

    def load_state_dict(self, sd):
        self.flat.copy_(sd['flat'].to(self.device))
        for s, t in zip(self.slots, sd['slots']):
            s.copy_(t.to(self.device))
        self.opt_steps = sd['opt_steps']
        self.iteration = sd['iteration']
