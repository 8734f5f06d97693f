"""A lane-by-lane Python model of fec.cu's warp-parallel RS(255,223) decoder and of its
position-wise re-encoder (rs_decode_warp, reencode_count_errors, build_rs_parity_map).
It exists so that the formulation can be checked against the oracle on the CPU, where no
GPU is available; tests/test_gpu_fec.py checks the CUDA code itself."""
import numpy as np

NN, NROOTS, FCR, PRIM, IPRIM, RSPAD = 255, 32, 112, 11, 116, 95
ROWS, COLS, SYMS, NBITS = 80, 65, 5200, 2566
CPOLYA, CPOLYB = 0x4F, 0x6D


class GF:
    def __init__(self, alpha_to, index_of):
        self.exp2 = np.array([alpha_to[i % 255] for i in range(510)] + [0, 0], dtype=np.int64)
        self.lg = np.array(index_of, dtype=np.int64)

    def mul(self, a, b):
        return int(self.exp2[self.lg[a] + self.lg[b]]) if (a and b) else 0

    def mul_exp(self, a, e):
        return int(self.exp2[self.lg[a] + e]) if a else 0

    def div(self, a, b):
        return int(self.exp2[self.lg[a] + 255 - self.lg[b]]) if a else 0


def rs_decode_warp(cw, gf):
    """cw: list of 255 ints (modified in place).  Returns 0 / count / -1 like the kernel."""
    syn = [0] * 32
    for lane in range(32):
        be = ((FCR + lane) * PRIM) % NN
        s = 0
        for j in range(RSPAD, NN):
            s = cw[j] ^ gf.mul_exp(s, be)
        syn[lane] = s
    if not any(syn):
        return 0
    lam = [0] * 32          # lam[l] = lambda[l+1]
    bq = [1] + [0] * 31     # bq[l] = b[l]
    L = 0
    for r in range(1, NROOTS + 1):
        d = syn[r - 1]
        for lane in range(32):
            if lane <= r - 2:
                d ^= gf.mul(lam[lane], syn[r - 2 - lane])
        b_up = [0] + bq[:-1]
        if d == 0:
            bq = b_up
        else:
            l_up = [1] + lam[:-1]
            lam_new = [lam[l] ^ gf.mul(d, bq[l]) for l in range(32)]
            if 2 * L <= r - 1:
                L = r - L
                bq = [gf.div(l_up[l], d) for l in range(32)]
            else:
                bq = b_up
            lam = lam_new
    deg = max([l + 1 for l in range(32) if lam[l]], default=0)
    lamf = [1] + lam
    roots, locs = [], []
    for i in range(1, NN + 1):
        q = 1
        for j in range(1, deg + 1):
            q ^= gf.mul_exp(lamf[j], (j * i) % NN)
        if q == 0:
            roots.append(i)
            locs.append((IPRIM * i - 1) % NN)
    if len(roots) != deg:
        return -1
    om = [0] * 32
    for lane in range(32):
        for j in range(0, min(deg, lane) + 1):
            om[lane] ^= gf.mul(syn[lane - j], lamf[j])
    bad = False
    for lane, (rt, loc) in enumerate(zip(roots, locs)):
        num1 = 0
        for i in range(NROOTS):
            num1 ^= gf.mul_exp(om[i], (i * rt) % NN)
        den = 0
        for i in range(0, (min(deg, NROOTS - 1) & ~1) + 1, 2):
            den ^= gf.mul_exp(lamf[i + 1], (i * rt) % NN)
        if den == 0:
            bad = True
        elif num1:
            cw[loc] ^= gf.div(gf.mul_exp(num1, (rt * (FCR - 1)) % NN), den)
    return -1 if bad else len(roots)


def rs_parity_map(alpha_to, index_of, rs_poly):
    gf = GF(alpha_to, index_of)
    g = [0] * 33
    for k in range(1, 17):
        g[k] = alpha_to[rs_poly[k - 1]]
    for k in range(17, 32):
        g[k] = g[32 - k]
    g[32] = 1
    reg = [g[k + 1] for k in range(32)]
    par = [None] * 128
    for i in range(127, -1, -1):
        par[i] = list(reg)
        fb = reg[0]
        reg = [reg[k + 1] ^ gf.mul(fb, g[k + 1]) for k in range(31)] + [fb]
    return par


def rs_parity(data128, par, gf):
    """The 32 parity bytes of one RS(160,128) block (lowest-order first), by the linear map."""
    out = []
    for k in range(32):
        p = 0
        for i in range(128):
            p ^= gf.mul(int(data128[i]), par[i][k])
        out.append(p)
    return out


def symbols_from_blocks(blocks, scrambler, sync):
    """The 5200 channel symbols (0/1) of a frame whose two RS blocks are given as 160 bytes each
    (128 data + 32 parity) — NOT necessarily code words: the frame-stage tests inject symbol
    errors here, behind the convolutional code.  Byte order of the stream the convolutional
    encoder sees (FECDecoder.java:614-671): data byte i = blocks[i & 1][i >> 1], then parity byte
    256 + 2k + r = blocks[r][128 + k]; scrambled; position by position as in fec.cu."""
    enc = [0] * 324
    for i in range(256):
        enc[i] = int(blocks[i & 1][i >> 1]) ^ scrambler[i]
    for k in range(32):
        for r in range(2):
            enc[256 + 2 * k + r] = int(blocks[r][128 + k]) ^ scrambler[256 + 2 * k + r]
    sym = np.zeros(SYMS, dtype=np.uint8)
    for p in range(SYMS):
        row, col = divmod(p, ROWS)
        if col == 0:
            sym[p] = 1 if sync[row] > 0 else 0
        else:
            k = col * COLS + row - COLS
            if k < 2 * NBITS:
                n = k >> 1
                byte = n >> 3
                two = ((enc[byte - 1] if byte else 0) << 8) | enc[byte]
                win = (two >> (7 - (n & 7))) & 0x7F
                sym[p] = (1 - (bin(win & CPOLYB).count("1") & 1)) if (k & 1) else (bin(win & CPOLYA).count("1") & 1)
    return sym


def reencode_symbols(out, par, gf, scrambler, sync):
    """The 5200 symbols (0/1) encode_FEC40 would produce for 256 data bytes, position by position."""
    blocks = []
    for r in range(2):
        d = [int(out[2 * i + r]) for i in range(128)]
        blocks.append(d + rs_parity(d, par, gf))
    return symbols_from_blocks(blocks, scrambler, sync)
