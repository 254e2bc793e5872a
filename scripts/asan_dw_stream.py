"""AddressSanitizer run of the depthwise row-stream lane program on the host (csrc/dwconv_stream.cu built with
-DMNB_DW_STREAM_EMUL -fsanitize=address): every tensor is an exactly-sized malloc block, so any out-of-bounds access of
the lane program (padding, edge strips, prefetch offsets) aborts.  Usage:

    ASAN=$(gcc -print-file-name=libasan.so)
    nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=address \
         -DMNB_DW_STREAM_EMUL -shared mnasnet-pytorch_b200/csrc/dwconv_stream.cu -o gpurun_out/asan/libdw_stream_emul.so
    LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0 python scripts/asan_dw_stream.py gpurun_out/asan/libdw_stream_emul.so
Last run (round 1): 17 shapes x (prefetch depth 1/2/3, 4- and 8-column strips) x fwd/dgrad/wgrad, clean."""
import ctypes
import sys

lib = ctypes.CDLL(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/asan/libdw_stream_emul.so")
fn = lib.mnb_emul_dw_stream
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 9 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
lib.mnb_emul_dw_stream_set_pd.argtypes = [ctypes.c_int]
lib.mnb_emul_dw_stream_set_tw8.argtypes = [ctypes.c_int]
libc = ctypes.CDLL("libc.so.6")
libc.malloc.restype = ctypes.c_void_p
libc.malloc.argtypes = [ctypes.c_size_t]
libc.free.argtypes = [ctypes.c_void_p]
libc.memset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]

def buf(nbytes, fill=0x3c):
    p = libc.malloc(nbytes)
    libc.memset(p, fill, nbytes)      # bf16 0x3c3c = 0.0115, float 0x3c3c3c3c = 0.0115: finite values everywhere
    return p

CASES = [(2, 12, 10, 32, 3, 8), (1, 9, 11, 72, 5, 12), (2, 7, 7, 48, 3, 6), (1, 30, 9, 48, 5, 3), (1, 6, 5, 240, 5, 8),
         (1, 33, 17, 16, 3, 40), (3, 14, 14, 64, 5, 5), (1, 5, 6, 1152, 3, 36), (2, 2, 2, 1152, 3, 592),
         (2, 4, 4, 576, 5, 444), (1, 1, 1, 240, 5, 444), (2, 64, 48, 72, 5, 444), (1, 9, 70, 32, 3, 16),
         (2, 20, 100, 48, 5, 24), (1, 12, 66, 16, 3, 10), (1, 3, 200, 16, 5, 7), (2, 57, 3, 24, 3, 9)]
for pd, tw8 in ((1, 0), (2, 0), (3, 0), (1, 1), (3, 1)):
    lib.mnb_emul_dw_stream_set_pd(pd)
    lib.mnb_emul_dw_stream_set_tw8(tw8)        # 8-column strips (3x3 forward / backward-data)
    for (N, H, W, C, k, warps) in CASES:
        n = N * H * W * C
        x, dz, out = buf(2 * n), buf(2 * n), buf(2 * n)
        s, t, b = buf(4 * C), buf(4 * C), buf(4 * C)
        w, dw = buf(4 * C * k * k), buf(4 * C * k * k)
        st = buf(8 * 2 * C, 0)
        for mode in (0, 1, 2):
            rc = fn(mode, x, s, t, w, b, dz, out, dw, st, N, H, W, C, k, warps, None)
            assert rc == 0, (rc, N, H, W, C, k)
        for p in (x, dz, out, s, t, b, w, dw, st):
            libc.free(p)
        print("ok", pd, tw8, (N, H, W, C, k, warps), flush=True)
print("ASAN CLEAN")
