"""Whole-step CUDA graph -- opt-in (`bench.py --graph`).

Why: a train step is ~900 kernel launches.  At the recipe's per-GPU batch on 8 GPUs (8 utterances x 500 frames) the
device needs ~3 ms for them but the host needs ~12 ms to issue them (measured: `profiles/bench_r1b_b8.json`,
12.7 ms/step at 8 utterances vs 22.7 ms at 64), i.e. the step is launch-bound.  Capturing `trainer._train_core`
(everything between the batch and the packed loss vector: forwards, backwards, weight-norm refreshes, gradient
all-reduces, Adam) once and replaying it removes the per-launch host cost.

What makes the step capturable (all of it already true in the eager path, or switched on here):
  * no host synchronisation inside `_train_core` (losses stay device scalars; `_parse_loss` fetches them once);
  * the Adam step count lives in device memory (`FusedAdam(capturable=True)` -> `crk_adam_step_dev`);
  * the library allocates nothing and never synchronises; torch's allocations go to the graph's private pool;
  * dropout masks come from torch's graph-safe Philox generator.
What it does not cover (falls back to eager, loudly): trainers with host-side randomness inside the step
(`cyclegan`'s random fake branch, `stargan` with `switch_update`), and any change of the schedule flags
(`gan_flag`, `cycle_flag`, `stop_generator`) or of a learning rate re-captures (they are baked into the launches).

Usage:
    step = GraphedTrainStep(trainer)            # trainer built as usual; optimizers switched to capturable
    values = step(batch)                        # dict of python floats, like trainer.train(batch, "train")
"""
import torch


class GraphedTrainStep:
    WARMUP = 3      # eager steps before capture (cuFuncSetAttribute calls, allocator warm-up, optimizer state creation)

    def __init__(self, trainer, phase="train"):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedTrainStep needs CUDA (there is no CPU path)")
        kind = type(trainer).__name__
        if kind == "CycleGANTrainer" or (kind == "StarGANTrainer" and trainer.conf.get("switch_update")):
            raise NotImplementedError(f"{kind} draws host-side random choices inside the step: not capturable")
        self.trainer, self.phase = trainer, phase
        for opt in trainer.optimizer.values():
            if not hasattr(opt, "capturable"):
                raise NotImplementedError("whole-step capture needs crank_b200's FusedAdam (device-side step count)")
            opt.capturable = True
        self._graphs = {}        # signature -> (graph, static batch, loss keys, packed loss tensor)
        self._eager_steps = 0

    def _signature(self, batch):
        t = self.trainer
        flags = (getattr(t, "gan_flag", False), getattr(t, "cycle_flag", False), getattr(t, "stop_generator", False))
        lrs = tuple(float(g["lr"]) for o in t.optimizer.values() for g in o.param_groups)
        shapes = tuple((k, tuple(v.shape), str(v.dtype)) for k, v in sorted(batch.items()) if isinstance(v, torch.Tensor))
        return flags, lrs, shapes

    def _invalidate_weight_caches(self):
        # replays update parameters without bumping their autograd versions: make every weight-norm cache
        # (parallel_wavegan.models._PackedConvNet.effective_weights) recompute on its next EAGER use
        for m in self.trainer.model.values():
            for p in m.parameters():
                torch.autograd.graph.increment_version(p)

    def _drop_weight_caches(self):
        # BEFORE a capture: forget every cached effective-weight buffer, so that each network's first use inside
        # the capture records its weight-norm kernel into a graph-pool buffer.  Without this the captured kernels
        # of the first G / D uses would read the buffer the preceding eager step filled, which no replay refreshes.
        for m in self.trainer.model.values():
            for sub in m.modules():
                if hasattr(sub, "_weff_key"):
                    sub._weff, sub._weff_key = None, None

    def __call__(self, batch):
        t = self.trainer
        if self._eager_steps < self.WARMUP:
            self._eager_steps += 1
            return t.train(batch, self.phase)
        sig = self._signature(batch)
        entry = self._graphs.get(sig)
        if entry is None:
            static = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # THIS call's step runs eagerly, on a side stream
                eager_values = t._parse_loss(t._train_core(static, self.phase))     # (torch's capture warm-up rule)
            torch.cuda.current_stream().wait_stream(side)
            from . import _dp

            _dp.flush()          # data parallel: no completion event of the eager steps' side-stream work may be waited on inside the capture
            self._drop_weight_caches()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = t._train_core(static, self.phase)
                keys = [k for k, v in loss.items() if isinstance(v, torch.Tensor)]
                packed = torch.stack([loss[k].detach().reshape(()).float() for k in keys])
            const = {k: float(v) for k, v in loss.items() if not isinstance(v, torch.Tensor)}
            self._graphs = {sig: (graph, static, keys, packed, const)}     # a schedule / lr change evicts the old graph + pool
            self._drop_weight_caches()
            self._invalidate_weight_caches()
            return eager_values                                 # the capture itself executed nothing
        graph, static, keys, packed, const = entry
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                static[k].copy_(v, non_blocking=True)
            else:
                static[k] = v
        graph.replay()
        self._invalidate_weight_caches()
        from . import _dp

        values = t._get_loss_dict()
        values.update(const)
        for k, v in zip(keys, _dp.average_loss_vector(packed.clone()).tolist()):
            values[k] = v
        t._last_loss_values = values
        if hasattr(t, "_flush_writer"):
            t._flush_writer({k: packed[i] for i, k in enumerate(keys)}, self.phase)
        return values
