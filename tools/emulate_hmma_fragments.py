"""CPU check of K5h's shared-memory addressing (wc_search.cu, wc_dist_topk_f16_kernel): emulates the TMA SWIZZLE_128B tile
image, the kernel's ldmatrix.x4 lane addresses and the m16n8k16 fragment layouts of the PTX ISA, and verifies that every
lane's accumulators hold A.B^T at the (row, column) the epilogue assumes.  Prints one line per consumer warp."""
import numpy as np
rng = np.random.default_rng(0)
BM=128; BKH=64
A = rng.integers(-3,4,size=(BM,BKH)).astype(np.float64)   # tile rows x k (halves)
B = rng.integers(-3,4,size=(BM,BKH)).astype(np.float64)
def swz(tile):
    img = np.zeros(BM*BKH)    # in halves; row r, chunk c (8 halves) at r*64 + ((c ^ (r&7))*8)
    for r in range(BM):
        for c in range(8):
            img[r*64 + ((c ^ (r&7))*8): r*64 + ((c ^ (r&7))*8)+8] = tile[r, c*8:(c+1)*8]
    return img
imgA, imgB = swz(A), swz(B)
def ldsm_x4(img, byte_addr_per_lane):
    # returns regs[lane][4] each a pair of halves
    regs = np.zeros((32,4,2))
    for m in range(4):
        rows = [img[(byte_addr_per_lane[m*8 + r]//2):(byte_addr_per_lane[m*8 + r]//2)+8] for r in range(8)]
        for t in range(32):
            regs[t, m] = rows[t//4][2*(t%4):2*(t%4)+2]
    return regs
for warp in range(8):
    acc = np.zeros((32,16,4))
    for ks in range(4):
        a_addr = []; 
        for lane in range(32):
            a_row = warp*16 + (lane & 7) + ((lane >> 3) & 1)*8
            a_kc = lane >> 4
            xr = lane & 7
            a_addr.append(a_row*128 + (((2*ks + a_kc) ^ xr) << 4))
        fa = ldsm_x4(imgA, a_addr)
        for np_ in range(8):
            b_addr = []
            for lane in range(32):
                b_row = ((lane >> 4) & 1)*8 + (lane & 7)
                b_kc = (lane >> 3) & 1
                xr = lane & 7
                b_addr.append(b_row*128 + np_*2048 + (((2*ks + b_kc) ^ xr) << 4))
            fb = ldsm_x4(imgB, b_addr)
            for half, (r0, r1) in enumerate(((0,1),(2,3))):
                nt = 2*np_ + half
                # emulate mma: build A(16x16), B(16x8) from fragments
                Am = np.zeros((16,16)); Bm = np.zeros((16,8))
                for t in range(32):
                    g, q = t//4, t%4
                    Am[g, 2*q:2*q+2] = fa[t,0]; Am[g+8, 2*q:2*q+2] = fa[t,1]
                    Am[g, 2*q+8:2*q+10] = fa[t,2]; Am[g+8, 2*q+8:2*q+10] = fa[t,3]
                    Bm[2*q:2*q+2, g] = fb[t,r0]; Bm[2*q+8:2*q+10, g] = fb[t,r1]
                Cm = Am @ Bm
                for t in range(32):
                    g, q = t//4, t%4
                    acc[t,nt,0] += Cm[g,2*q]; acc[t,nt,1] += Cm[g,2*q+1]; acc[t,nt,2] += Cm[g+8,2*q]; acc[t,nt,3] += Cm[g+8,2*q+1]
    # check against A B^T for the warp's rows
    want = A[warp*16:(warp+1)*16] @ B.T     # 16 x 128
    ok = True
    for t in range(32):
        g, q = t//4, t%4
        for nt in range(16):
            for hh in range(2):
                for e in range(2):
                    row = hh*8 + g; col = nt*8 + 2*q + e
                    if acc[t,nt,2*hh+e] != want[row,col]: ok = False
    print("warp", warp, "ok" if ok else "MISMATCH")
