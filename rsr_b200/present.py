"""The presenting GPU's frame buffer, shared with the other ranks of the node.

Split-frame rendering (BASELINE.json configs[4]): every rank renders the sub-frames it
owns; the only exchange is getting resolved 8-bit pixels to the GPU that presents.  Instead of
rendering into a local buffer and gathering afterwards, rank 0 allocates the whole frame once and
hands the other ranks a CUDA-IPC mapping of it; their tile kernels then resolve straight into it --
the pixels cross NVLink / NVSwitch as peer stores while the kernel is still rasterising other
tiles, and a frame ends with one small barrier instead of a gather plus an assembly copy.
PyTorch is plumbing here (allocation, IPC handle exchange over the process group).
"""
from __future__ import annotations


class PresentedFrame:
    """frame: (buffers, height, width) int32 tensor on the presenting rank's GPU (double buffered presentation: buffers = 2);
    on other ranks a peer mapping.
    counter: one int64 per rank next to it (same sharing): rank r adds 1 to counter r behind its sub-frames of a frame
    (rsrcu_signal_counter); streams wait on all of them (rsrcu_wait_counters)."""

    def __init__(self, width: int, height: int, rank: int, local_rank: int, world: int, dist=None, presenter: int = 0, buffers: int = 1):
        import torch
        from torch.multiprocessing.reductions import reduce_tensor
        self.width, self.height = width, height
        self.rank, self.world, self.presenter = rank, world, presenter
        self.local = None
        self.local_counter = None
        if rank == presenter:
            self.local = torch.zeros((buffers, height, width), dtype=torch.int32, device=f"cuda:{local_rank}")
            self.local_counter = torch.zeros(max(world, 1), dtype=torch.int64, device=f"cuda:{local_rank}")
        if world == 1:
            self.frame, self.counter = self.local, self.local_counter
            self.presenter_device = local_rank
            return
        box = [None]
        if rank == presenter:
            box = [(reduce_tensor(self.local), reduce_tensor(self.local_counter), local_rank)]
        dist.broadcast_object_list(box, src=presenter)
        (fn, args), (cfn, cargs), self.presenter_device = box[0]
        if rank == presenter:
            self.frame, self.counter = self.local, self.local_counter
        else:
            # open the handles from THIS rank's device (argument 6 of rebuild_cuda_tensor = the device whose context maps
            # the memory): cudaIpcOpenMemHandle then maps the presenter's memory for this GPU with peer access
            # over NVLink.  The tensors only carry the addresses; they are never used for torch compute.
            args, cargs = list(args), list(cargs)
            args[6] = local_rank
            cargs[6] = local_rank
            self.frame = fn(*args)
            self.counter = cfn(*cargs)

    def pointer(self, x0: int, y0: int, buffer: int = 0) -> int:
        """device address of pixel (x0, y0) of frame buffer `buffer` -- valid in this process for kernels on any GPU with peer access"""
        return self.frame.data_ptr() + 4 * ((buffer * self.height + y0) * self.width + x0)

    def counter_pointer(self, rank: int = 0) -> int:
        """device address of rank `rank`'s completion counter (the counters are consecutive: wait on counter_pointer(0), world)"""
        return self.counter.data_ptr() + 8 * rank

    @property
    def stride_px(self) -> int:
        return self.width
