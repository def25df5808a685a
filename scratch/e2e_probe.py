import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import kmersgwas_b200 as kg
import bench
n, p, R = 1135, 101, 1 << 22
stride = 19
y = bench.phenotypes(n, p)
idx = np.arange(n)
mw, mb = (idx // 64).astype(np.uint32), (idx % 64).astype(np.uint32)
stream = torch.cuda.Stream()
sess = kg.Session(n, mw, mb, y, 57, 10001, stream=stream.cuda_stream)
abi = kg.load(); h = sess.ctx_handle
prefill = int(sys.argv[1]) if len(sys.argv) > 1 else 0
with torch.cuda.stream(stream):
    bufs = [torch.empty(R * stride, dtype=torch.int64, device='cuda') for _ in range(2)]
    host = [torch.empty(R * stride, dtype=torch.int64).pin_memory() for _ in range(3)]
    # plain H2D bandwidth
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(6): bufs[i % 2].copy_(host[i % 3], non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print('plain H2D GB/s', 6 * R * stride * 8 / (t1 - t0) / 1e9)
    step = 0
    for ps in range(prefill + 3):
        abi.kg_synth_rows_device(h, 1, step * R, R, bufs[ps % 2].data_ptr())
        sess.associate(bufs[ps % 2].data_ptr(), R, step * R); step += 1
    sess.finish()
    for i in range(3):
        abi.kg_synth_rows_device(h, 1, (step + i) * R, R, bufs[0].data_ptr()); stream.synchronize(); host[i].copy_(bufs[0])
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(10):
            sess.associate(host[i % 3].data_ptr(), R, (step + i % 3) * R)
        sess.finish(); torch.cuda.synchronize()
        t1 = time.perf_counter()
        print('prefill steps', prefill, 'e2e ms/step', (t1 - t0) * 100, 'GB/s', 10 * R * stride * 8 / (t1 - t0) / 1e9, sess.stats())
