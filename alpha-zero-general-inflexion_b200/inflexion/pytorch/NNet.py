"""``NNetWrapper``: the reference's network wrapper API (inflexion/pytorch/NNet.py:27-120) -- ``train``,
``predict``, ``save_checkpoint``, ``load_checkpoint`` with the same hyper-parameters, losses and file format --
plus ``predict_batch`` for the batched MCTS and data-parallel training when ``torch.distributed`` is
initialised (gradients all-reduced over NCCL/NVLink; nothing else is communicated).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.optim as optim

from ...NeuralNet import NeuralNet
from ...utils import AverageMeter, dotdict
from .InflexionNNet import FoldedInference, InflexionNNet

# hyper-parameters of the reference (NNet.py:17-24)
args = dotdict({
    "lr": 0.001,
    "dropout": 0.3,
    "epochs": 10,
    "batch_size": 512,
    "cuda": torch.cuda.is_available(),
    "num_channels": 512,
})


class NNetWrapper(NeuralNet):
    def __init__(self, game, device=None, infer_dtype=torch.float16):
        self.nnet = InflexionNNet(game, args)
        self.depth, self.board_x, self.board_y = 4, *game.board_shape
        self.action_size = game.max_actions
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if args.cuda else torch.device("cpu")
        self.device = torch.device(device)
        self.nnet.to(self.device)
        self.infer_dtype = infer_dtype
        self._folded = None            # FoldedInference, rebuilt after every weight change
        self.weights_version = 0       # bumped with it: owners of CUDA graphs that captured the folded net re-capture

    # ---- training (NNet.py:36-76) -----------------------------------------------------------------
    def train(self, examples, verbose=True):
        """examples: list of (planes, pi, v).  Adam with default lr (the reference never passes args.lr,
        NNet.py:40), ``epochs`` x ``len/batch`` steps, each batch drawn WITH replacement (NNet.py:52).
        Under torch.distributed the GLOBAL batch stays ``batch_size`` (each rank draws batch_size/world rows from its own
        examples), gradients are averaged by a bucketed all-reduce overlapped with backward, and on GPUs BatchNorm uses
        the statistics of the global batch -- one data-parallel step is the reference's step on the union of the
        ranks' draws."""
        dist = torch.distributed if torch.distributed.is_available() and torch.distributed.is_initialized() else None
        world = dist.get_world_size() if dist else 1
        # (the reference's optimizer, NNet.py:40; on a GPU the fused implementation: one kernel per step instead of a dozen)
        optimizer = optim.Adam(self.nnet.parameters(), fused=True) if self.device.type == "cuda" else optim.Adam(self.nnet.parameters())
        # ``examples`` is the reference's list of (planes, pi, v), or an on-demand set (selfplay.ExampleSet /
        # ExampleUnion: ``len`` + ``batch(ids)``) that expands only the rows a batch draws
        on_demand = hasattr(examples, "batch")
        if not on_demand:
            boards_all = torch.as_tensor(np.array([e[0] for e in examples]).astype(np.float32))
            pis_all = torch.as_tensor(np.array([e[1] for e in examples], dtype=np.float32))
            vs_all = torch.as_tensor(np.array([e[2] for e in examples], dtype=np.float32))
        # steps per epoch: one pass over the examples (NNet.py:50).  Under torch.distributed the examples are
        # rank-local and their number differs from rank to rank (games end at different moves), so every rank must
        # agree on the count or the gradient all-reduce deadlocks: one pass over the GLOBAL set at a global batch of
        # world x batch_size.
        n_total = len(examples)
        per_rank_batch = args.batch_size
        if dist and world > 1:
            # (i) every rank must enter with the SAME net: parameters and buffers are broadcast from rank 0 (nothing else
            # guarantees it -- NNetWrapper.__init__ is randomly initialised per process);
            for t in list(self.nnet.parameters()) + list(self.nnet.buffers()):
                dist.broadcast(t.data, 0)
            # gradients: bucketed all-reduce (mean) overlapped with the backward pass (DistributedDataParallel: NCCL over
            # NVLink on GPUs, gloo in the CPU tests).  On GPUs the BatchNorm layers become SyncBatchNorm for the length of
            # train(), so batch statistics are those of the GLOBAL batch, as in the reference's single process
            # (NNet.py:52,58); the original modules (sharing the same tensors) are put back at the end.
            from torch.nn.parallel import DistributedDataParallel as DDP
            swapped = {}
            # ``args.dp_sync_bn = False`` (optional): every rank normalises with the statistics of its OWN 512/world rows --
            # no collective inside the layers (SyncBatchNorm costs 12 small all-gathers / all-reduces per step, about as
            # much as the rest of a latency-bound step at 8 GPUs), at the price of leaving the reference's semantics;
            # the running statistics are then averaged over the ranks at the end of train().
            if self.device.type == "cuda" and args.get("dp_sync_bn", True):
                for name, mod in list(self.nnet.named_children()):
                    if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                        swapped[name] = mod
                        setattr(self.nnet, name, torch.nn.SyncBatchNorm.convert_sync_batchnorm(mod))
            model = DDP(self.nnet, device_ids=[self.device.index] if self.device.type == "cuda" else None,
                        broadcast_buffers=False, gradient_as_bucket_view=True, bucket_cap_mb=8)
            # (ii) the step count and the "a rank has no examples" error are decided COLLECTIVELY, so that all ranks
            # raise together instead of one raising while the others hang in the first gradient all-reduce;
            t = torch.tensor([n_total, -n_total], dtype=torch.int64, device=self.device)
            dist.all_reduce(t[:1])
            dist.all_reduce(t[1:], op=dist.ReduceOp.MAX)
            n_total, n_min = int(t[0]), -int(t[1])
            if n_min == 0:
                raise ValueError("NNetWrapper.train: a rank has no examples")
            # (iii) the reference's schedule is kept: GLOBAL batch = batch_size (NNet.py:21,52), i.e. batch_size/world
            # rows per rank, Adam's lr unchanged, steps = global examples / batch_size per epoch (NNet.py:50).
            # ``args.dp_global_batch`` (optional) overrides the global batch for runs that want larger steps.
            global_batch = int(args.get("dp_global_batch") or args.batch_size)
            per_rank_batch = max(1, global_batch // world)
            steps = int(n_total / (per_rank_batch * world))
        else:
            steps = int(n_total / args.batch_size)
            model, swapped = self.nnet, {}
        history = []
        for epoch in range(args.epochs):
            if verbose:
                print("EPOCH ::: " + str(epoch + 1))
            self.nnet.train()
            pi_losses, v_losses = AverageMeter(), AverageMeter()
            # Batches are prepared a chunk of steps at a time (one index draw, one expansion / one host-to-device copy
            # for the whole chunk, sliced per step) and the loss meters are read once per epoch: no per-step launch
            # storm and no per-step host sync.  The draws are the same uniform-with-replacement samples (NNet.py:52).
            chunk = 16
            loss_sum = torch.zeros(2, dtype=torch.float64, device=self.device)
            done = 0
            while done < steps:
                k = min(chunk, steps - done)
                ids = np.random.randint(len(examples), size=k * per_rank_batch)
                if on_demand:
                    boards_c, pis_c, vs_c = (t.to(self.device) for t in examples.batch(ids))
                else:
                    ids_t = torch.from_numpy(ids)
                    boards_c = boards_all[ids_t].to(self.device, non_blocking=True)
                    pis_c = pis_all[ids_t].to(self.device, non_blocking=True)
                    vs_c = vs_all[ids_t].to(self.device, non_blocking=True)
                for j in range(k):
                    sl = slice(j * per_rank_batch, (j + 1) * per_rank_batch)
                    boards, target_pis, target_vs = boards_c[sl], pis_c[sl], vs_c[sl]
                    out_pi, out_v = model(boards)
                    l_pi = self.loss_pi(target_pis, out_pi)
                    l_v = self.loss_v(target_vs, out_v)
                    optimizer.zero_grad(set_to_none=False)
                    (l_pi + l_v).backward()                  # (DDP: the gradient buckets are all-reduced while this runs)
                    optimizer.step()
                    loss_sum += torch.stack([l_pi.detach(), l_v.detach()]).to(torch.float64)
                done += k
            if steps:
                lp, lv = (loss_sum / steps).tolist()
                pi_losses.update(lp, steps * per_rank_batch)
                v_losses.update(lv, steps * per_rank_batch)
            history.append((pi_losses.avg, v_losses.avg))
            if verbose:
                print(f"  Loss_pi={pi_losses} Loss_v={v_losses}")
        if dist and world > 1:
            # BatchNorm running statistics are computed from rank-local batches: average them so that every rank
            # leaves train() with the SAME net (parameters already agree through the all-reduced gradients) -- the
            # ranks play the next iteration's games with it and rank 0 writes the checkpoint for all of them
            # (with SyncBatchNorm they already agree; averaging identical values is a no-op up to rounding)
            if not swapped:
                for name, buf in self.nnet.named_buffers():
                    if buf.dtype.is_floating_point:
                        dist.all_reduce(buf)
                        buf /= world
            for name, mod in swapped.items():              # plain BatchNorm modules back (they share the trained tensors)
                setattr(self.nnet, name, mod)
            del model
        self.weights_changed()
        return history

    # ---- inference --------------------------------------------------------------------------------
    def predict(self, board):
        """Batch-1 forward in fp32, exactly the reference's call (NNet.py:78-94): returns
        (float32[343], float32 ndarray of shape (1,))."""
        x = torch.as_tensor(np.asarray(board).astype(np.float32)).to(self.device).view(1, self.depth, self.board_x, self.board_y)
        self.nnet.eval()
        with torch.no_grad():
            log_pi, v = self.nnet(x)
        return torch.exp(log_pi).cpu().numpy()[0], v.cpu().numpy()[0]

    def predict_batch(self, planes: torch.Tensor):
        """float32 [B,4,7,7] on the device -> (pi float32 [B,343], v float32 [B]).  BN-folded 16-bit graph
        (fp16 by default: see FoldedInference for the measured deviation from the fp32 module)."""
        if self._folded is None:
            self.nnet.eval()
            self._folded = FoldedInference(self.nnet, self.infer_dtype)
        return self._folded(planes)

    def predict_head(self, planes: torch.Tensor, out: torch.Tensor):
        """float32 [B,4,7,7] -> raw head output written into ``out`` (float16 [B,344]: 343 policy logits + the value
        pre-activation).  For ``SelfPlayEngine(fused_head=True)``: the tree kernel applies softmax and tanh."""
        if self._folded is None:
            self.nnet.eval()
            self._folded = FoldedInference(self.nnet, self.infer_dtype)
        if self._folded.dtype != torch.float16:
            raise ValueError("predict_head needs float16 inference (the kernel reads binary16 rows)")
        return self._folded.head_into(planes, out)

    def weights_changed(self):
        self._folded = None
        self.weights_version += 1

    # ---- losses (NNet.py:96-100) --------------------------------------------------------------------
    def loss_pi(self, targets, outputs):
        return -torch.sum(targets * outputs) / targets.size()[0]

    def loss_v(self, targets, outputs):
        return torch.sum((targets - outputs.view(-1)) ** 2) / targets.size()[0]

    # ---- checkpoints: torch.save({'state_dict': ...}) (NNet.py:102-120) --------------------------------
    def save_checkpoint(self, folder="checkpoint", filename="checkpoint.pth.tar"):
        os.makedirs(folder, exist_ok=True)
        torch.save({"state_dict": self.nnet.state_dict()}, os.path.join(folder, filename))

    def load_checkpoint(self, folder="checkpoint", filename="checkpoint.pth.tar"):
        filepath = os.path.join(folder, filename)
        if not os.path.exists(filepath):
            raise FileNotFoundError(f"No model in path {filepath}")
        checkpoint = torch.load(filepath, map_location=self.device)
        self.nnet.load_state_dict(checkpoint["state_dict"])
        self.weights_changed()
