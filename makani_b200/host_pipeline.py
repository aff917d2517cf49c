"""Host <-> device staging for a training loop around the spectral block.

The SFNO block at 721x1440x73 moves 151 MB of input per sample over PCIe (~2.7 ms) for ~1 ms of GPU work, so a loop that copies,
computes and reads back in one stream spends three quarters of its time on the bus.  `HostFeed` is the usual remedy (the role of the
DALI pipeline's `prefetch_queue_depth=2` in the reference, /root/reference/makani/utils/dataloaders/data_loader_dali_2d.py:43): the
input of step i+1 is copied on a side stream into the other of two device buffers while step i computes, and results are read
back on a third stream.  Every step still pays its own copies; they overlap the previous / next step's kernels.

    feed = HostFeed(x_host.shape, x_host.dtype, device)
    feed.push(x_host)                      # input of step 0
    for i in range(steps):
        x = feed.pop()                     # compute stream waits for this step's input
        if i + 1 < steps:
            feed.push(next_host_batch)     # overlaps the kernels below
        y, _ = conv(x); y.backward(g)
        feed.release(x)                    # the buffer may be overwritten once the compute stream gets here
        feed.read_back(grad, grad_host)    # device -> pinned host on the read-back stream
    feed.drain()                           # compute stream waits for all outstanding read-backs
"""
import torch


class HostFeed:
    def __init__(self, shape, dtype, device, depth=2):
        self.device = torch.device(device)
        self.bufs = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(depth)]
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]   # copy of buffer b finished
        self.free = [None] * depth                               # compute no longer reads buffer b
        self.head = 0      # next buffer to fill
        self.tail = 0      # next buffer to hand out
        self.pending = 0

    def push(self, host_tensor):
        """Enqueue the copy of one pinned host batch into the next free device buffer (side stream)."""
        if self.pending >= len(self.bufs):
            raise RuntimeError("HostFeed.push: all device buffers are in flight; pop() / release() first")
        if not host_tensor.is_pinned():
            raise ValueError("HostFeed.push needs pinned host memory (tensor.pin_memory()) for an asynchronous copy")
        b = self.head
        with torch.cuda.stream(self.h2d):
            if self.free[b] is not None:
                self.h2d.wait_event(self.free[b])
            self.bufs[b].copy_(host_tensor, non_blocking=True)
            self.ready[b].record(self.h2d)
        self.head = (b + 1) % len(self.bufs)
        self.pending += 1

    def pop(self):
        """Device tensor of the oldest pushed batch; the current stream waits for its copy."""
        if self.pending == 0:
            raise RuntimeError("HostFeed.pop: nothing was pushed")
        b = self.tail
        torch.cuda.current_stream(self.device).wait_event(self.ready[b])
        self.tail = (b + 1) % len(self.bufs)
        self.pending -= 1
        return self.bufs[b]

    def release(self, x):
        """Mark the buffer behind `x` reusable from this point of the current stream on."""
        for b, buf in enumerate(self.bufs):
            if buf.data_ptr() == x.data_ptr():
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self.free[b] = ev
                return
        raise ValueError("HostFeed.release: tensor does not belong to this feed")

    def read_back(self, dev_tensor, host_tensor):
        """Copy a result to pinned host memory on the read-back stream, ordered after the current stream's work so far."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        dev_tensor.record_stream(self.d2h)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(ev)
            host_tensor.copy_(dev_tensor, non_blocking=True)

    def drain(self):
        """Make the current stream wait for every outstanding copy (call before timing stops / before reading host results)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.d2h)
        cur.wait_stream(self.h2d)
