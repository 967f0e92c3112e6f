"""Host <-> device plumbing for callers whose batches live in host memory (the reference's DataLoader path).

``HostStepRunner`` runs ``fn(device_batch) -> device_result`` for a stream of pinned host batches and lands every
result in pinned host memory.  Copies run on their own CUDA streams and are double-buffered, so the upload of
batch i+1 and the download of result i-1 overlap the kernels of batch i; nothing here synchronises the host
except ``result()`` / ``drain()``.
"""
from __future__ import annotations

import os

import torch


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int):
    """Pins the calling process to the CPU cores of the NUMA node its GPU hangs off (sysfs ``local_cpulist`` of the GPU's PCI
    function), so that the issuing thread, its pinned staging buffers (first touch) and the PCIe root port share a socket.
    With one process per GPU this keeps 8 ranks from piling onto node 0.  Returns a description, or None if the topology
    is not exposed (containers without sysfs) -- never raises."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        cpus = _parse_cpulist(open(base + "/local_cpulist").read())
        node = int(open(base + "/numa_node").read().strip())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        torch.set_num_threads(max(1, min(len(allowed), torch.get_num_threads())))
        return {"pci": bus, "numa_node": node, "cores": len(allowed), "first_core": allowed[0], "last_core": allowed[-1]}
    except Exception:
        return None


class HostStepRunner:
    def __init__(self, fn, in_shape, out_shape, device, depth: int = 2, dtype=torch.float32, graph: bool = False,
                 blocking_submit: bool = False):
        if torch.device(device).type != "cuda":
            raise RuntimeError("HostStepRunner needs a CUDA device (no CPU fallback)")
        self.fn, self.depth, self.device = fn, int(depth), torch.device(device)
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.dev_in = [torch.empty(in_shape, dtype=dtype, device=self.device) for _ in range(self.depth)]
        self.host_out = [torch.empty(out_shape, dtype=dtype).pin_memory() for _ in range(self.depth)]
        mk = lambda: [torch.cuda.Event() for _ in range(self.depth)]  # noqa: E731
        self.in_ready, self.in_free, self.out_ready, self.done = mk(), mk(), mk(), mk()
        self.step = 0
        # blocking_submit=True restores the old behaviour (submit waits on the host until the result slot it is about to reuse
        # has been written); by default ``result(i)`` is the only place that waits, and a result must be read before step
        # i + depth is submitted
        self.blocking_submit = bool(blocking_submit)
        # graph=True: fn is captured once per device buffer into a CUDA graph (after one eager call) and replayed, which
        # takes the Python issue cost of the module stack off the critical path.  fn must then be capture-safe: no host
        # synchronisation, no data-dependent Python control flow, parameters unchanged between steps.
        self.use_graph = bool(graph)
        self.graphs = [None] * self.depth
        self.static_out = [None] * self.depth
        self.h2d_bytes = self.dev_in[0].numel() * self.dev_in[0].element_size()
        self.d2h_bytes = self.host_out[0].numel() * self.host_out[0].element_size()

    def submit(self, host_batch: torch.Tensor) -> int:
        """Enqueues upload -> fn -> download for one pinned host batch; returns the step index."""
        i, s = self.step, self.step % self.depth
        compute = torch.cuda.current_stream(self.device)
        if i >= self.depth:
            self.h2d.wait_event(self.in_free[s])     # the kernels of step i-depth have consumed this device buffer
            # the download of step i-depth has read the (static) output this step's kernels will overwrite: ordered on the
            # DEVICE -- the issuing thread never blocks here, so it can queue steps as fast as Python allows
            compute.wait_event(self.done[s])
            if self.blocking_submit:
                self.done[s].synchronize()
        with torch.cuda.stream(self.h2d):
            self.dev_in[s].copy_(host_batch, non_blocking=True)
            self.in_ready[s].record(self.h2d)
        compute.wait_event(self.in_ready[s])
        if not self.use_graph:
            out = self.fn(self.dev_in[s])
        else:
            if self.graphs[s] is None:
                self.fn(self.dev_in[s])                  # eager warm-up (lazy initialisation, weight packing, allocator)
                compute.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.static_out[s] = self.fn(self.dev_in[s])
                self.graphs[s] = g
            self.graphs[s].replay()
            out = self.static_out[s]
        self.in_free[s].record(compute)
        self.out_ready[s].record(compute)
        self.d2h.wait_event(self.out_ready[s])
        with torch.cuda.stream(self.d2h):
            self.host_out[s].copy_(out.reshape(self.host_out[s].shape), non_blocking=True)
            self.done[s].record(self.d2h)
        if not self.use_graph:
            out.record_stream(self.d2h)
        self.step += 1
        return i

    def result(self, i: int) -> torch.Tensor:
        """Pinned host tensor holding the result of step i (valid until step i+depth is submitted)."""
        if not (self.step - self.depth <= i < self.step):
            raise IndexError("result %d is no longer (or not yet) buffered" % i)
        self.done[i % self.depth].synchronize()
        return self.host_out[i % self.depth]

    def drain(self) -> None:
        self.h2d.synchronize()
        self.d2h.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
