"""Copy-only ceiling of the host feed: every rank uploads the bytes of one bench step (65 1080p RGB24
frames, 404 MB) from page-locked host memory with NO kernels running, all ranks at once, timed with CUDA
events (max over ranks).  Variants: one copy stream / two copy streams / write-combined staging memory
(stb_host_alloc(.., write_combined=1)).  Run under torchrun with the rank count of interest:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/copy_ceiling.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from scannertools_b200 import _lib  # noqa: E402


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _lib.load()
    frames, fbytes = 65, 1080 * 1920 * 3
    nbytes = frames * fbytes
    dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')

    def host_buffer(wc):
        p = C.c_void_p()
        _lib.check(lib.stb_host_alloc(nbytes, 1 if wc else 0, C.byref(p)), lib)
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
        arr[:] = 7
        return p, torch.from_numpy(arr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters=8):
        fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) * 1e-3 / iters], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    out = {'n_gpus': world, 'bytes_per_gpu_and_step': nbytes}
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for name, wc in (('pinned', False), ('write_combined', True)):
        p, host = host_buffer(wc)
        assert host.is_pinned(), 'torch does not see the buffer as page-locked'

        def one():
            dev.copy_(host, non_blocking=True)

        def two():
            cur = torch.cuda.current_stream()
            half = nbytes // 2
            for i, (a, b) in enumerate(((0, half), (half, nbytes))):
                streams[i].wait_stream(cur)
                with torch.cuda.stream(streams[i]):
                    dev[a:b].copy_(host[a:b], non_blocking=True)
            for s in streams:
                cur.wait_stream(s)
        for sname, fn in (('one_stream', one), ('two_streams', two)):
            t = timed(fn)
            out['%s_%s' % (name, sname)] = {'gb_per_s_per_gpu': nbytes / t / 1e9, 'gb_per_s_all': world * nbytes / t / 1e9,
                                            'frames_per_s_all': world * 64 / t}
        del host
        lib.stb_host_free(p)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
