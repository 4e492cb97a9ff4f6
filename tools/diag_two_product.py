"""Two-product vs three-product exact filter at BASELINE config 3 (25 000 x 112 000 x 1024, top-50): time per search,
queries re-run under the wide bound (development helper)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
import importlib
S = importlib.import_module("b200knn.search")
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
g = b200knn.normalize(torch.randn((112000, 1024), generator=gen, device=dev))
q = b200knn.normalize(torch.randn((25000, 1024), generator=gen, device=dev))
ix = b200knn.FlatIndex(1024, "cosine", "fp32").adopt(g)
os.environ["KNN_EXACT_ENGINE"] = "tensor"
for products, slack in (("3", None), ("2", None)):
    os.environ["KNN_EXACT_PRODUCTS"] = products
    if slack:
        os.environ["KNN_TWO_PRODUCT_KC"] = str(slack)
    for _ in range(3):
        v, i = ix.search(q, 50)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        v, i = ix.search(q, 50)
    torch.cuda.synchronize()
    print(products, slack, f"{(time.perf_counter() - t0) / 5 * 1e3:.2f} ms", "rerun", S._search_exact_tensor.last_two_product_rerun,
          "ffma", S._search_exact_tensor.last_unverified, flush=True)
    if products == "3":
        ref = (v, i)
    else:
        print("  identical", torch.equal(v, ref[0]) and torch.equal(i, ref[1]))
# stage breakdown (CUDA events), two products: split of the queries, filter (kc = 128 / 64), re-scoring
L = importlib.import_module("b200knn._lib")
lib = L.load()
filt = ix._filter
qsq = S.row_sqnorm(q)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for dt, kc, name in ((L.KNN_BF16X2, 128, "x2 kc128"), (L.KNN_BF16X2, 64, "x2 kc64"), (L.KNN_BF16X3, 64, "x3 kc64"), (L.KNN_BF16X3, 128, "x3 kc128")):
    eps = S.filter_error_bound(qsq, filt.max_sqnorm, 1024, "cosine", filt.lo_max_sqnorm if dt == L.KNN_BF16X2 else None)
    for rep in range(3):
        e0 = ev()
        q3 = S.split_bf16x3(q, "queries")
        e1 = ev()
        cv = torch.empty((25000, kc), dtype=torch.float32, device=dev); ci = torch.empty((25000, kc), dtype=torch.int64, device=dev)
        nb = lib.knn_search_workspace(25000, 112000, q3.shape[1], dt, kc)
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        lib.knn_search(q3.data_ptr(), filt.split.data_ptr(), None, None, 25000, 112000, q3.shape[1], dt, kc, 0, 0, 0, 0,
                       cv.data_ptr(), ci.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        e2 = ev()
        ov = torch.empty((25000, 50), dtype=torch.float32, device=dev); oi = torch.empty((25000, 50), dtype=torch.int64, device=dev)
        fl = torch.empty((25000,), dtype=torch.int32, device=dev)
        lib.knn_rescore_exact(q.data_ptr(), g.data_ptr(), None, None, 25000, 112000, 1024, 0, 0, 0, 0, cv.data_ptr(), ci.data_ptr(),
                              kc, 50, eps.data_ptr(), ov.data_ptr(), oi.data_ptr(), fl.data_ptr(), torch.cuda.current_stream().cuda_stream)
        e3 = ev()
        torch.cuda.synchronize()
    print(name, f"split {e0.elapsed_time(e1):.2f} filter+merge {e1.elapsed_time(e2):.2f} rescore {e2.elapsed_time(e3):.2f} ms; unverified {int(fl.sum())}")
