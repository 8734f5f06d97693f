"""Handles give back what they took: create / use / destroy in a loop and the device's free memory
(cudaMemGetInfo through torch — plumbing only) ends where it started; the same for a context of
its own with its streams, events and pinned buffers."""
import numpy as np
import pytest

import jsdrcuda as J

pytestmark = pytest.mark.gpu


def free_bytes():
    import torch
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info(0)[0]


def one_round(ctx, rng, k):
    rate = 96000
    n = [4096, 9600, 65536, 32768, 1000][k % 5]
    nchan = [1, 40, 3][k % 3]
    adsc = J.AudioDescriptor(rate, blen=n * 4)
    autotune = k % 4 == 3 and n >= 1024                   # the transform length is the bank's block length
    f = J.fft(ctx, None, adsc, max_batch=nchan * 2, n=n)
    b = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=np.linspace(1000.0, 30000.0, nchan), max_block=n if autotune else 2 * n)
    if k % 2:
        b.enable_fec(np.zeros((2, 256), np.int16), max_frames=4)
    if autotune:
        b.set_autotune(True)
    d = J.demod(ctx, adsc, nchan=nchan, max_block=n)
    r = J.fir(ctx, float(rate), nchan=nchan, max_block=n)
    raw = rng.integers(-9000, 9000, (nchan, 4 * n)).astype(np.int16)
    psd = np.zeros((nchan * 2, n + 2), np.float32)
    if not autotune:
        J.pump_receive_s16(f, b, raw, 2, psd)
        pix = np.zeros((nchan * 2, 100), np.int32)
        J.pump_waterfall_s16(f, b, raw, 2, 100, pix, np.zeros((nchan * 2, 2), np.float32))
    else:
        b.receive_raw(np.ascontiguousarray(raw[:, :2 * n]), shared=False)   # auto-tune: whole blocks of max_block
    b.read_ds()
    d.weights(3000, 6000)
    d.receive_audio(rng.uniform(-1, 1, (nchan, 2 * n)).astype(np.float32))
    r.weights(1000, 2000)
    r.filter(rng.integers(-1000, 1000, (nchan, n)).astype(np.int32))
    pinned = ctx.host_alloc((1 << 20,), np.uint8)
    dev = ctx.dev_alloc(1 << 22)
    dev.free()
    ctx.host_free(pinned)
    for h in (f, b, d, r):
        h.close()


def test_handles_return_their_device_memory(ctx):
    rng = np.random.default_rng(1)
    for k in range(5):                                       # warm up: lazy module loading, allocator pools
        one_round(ctx, rng, k)
    before = free_bytes()
    for k in range(40):
        one_round(ctx, rng, k)
    after = free_bytes()
    assert before - after < (8 << 20), f"{(before - after) >> 20} MiB of device memory not returned after 40 rounds"


def test_contexts_return_everything():
    rng = np.random.default_rng(2)
    c = J.Context(0)
    one_round(c, rng, 0)
    c.close()
    before = free_bytes()
    for k in range(10):
        c = J.Context(0)
        one_round(c, rng, k)
        c.close()
    after = free_bytes()
    assert before - after < (8 << 20), f"{(before - after) >> 20} MiB of device memory not returned after 10 contexts"
