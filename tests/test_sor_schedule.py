"""Host-side model of k_sor_small's schedule (flowonthego_b200/csrc/varref.cu) against the lexicographic sweep of
sor_coupled (kroeger/FDF1.0.1/solver.c:77-421, as restated in oracle/dis_oracle.c).

The kernel runs all T sweeps of a level in one CTA: thread (s, j) owns image row j in sweep s and, at step t, updates
the C pixels i0 ... i0 + C - 1 of that row, i0 = C (t - j - 2 s), from
  * its own previous result (left neighbour) and the values it read one step earlier (the pixel's value before this
    sweep),
  * what the rows above / below and the previous sweep PUBLISHED at step t - 1 (a double-buffered array),
  * for sweep 0, the du field from before the launch (zero past the last column).
This file restates exactly that data flow in numpy scalars -- every operand is taken from where the kernel takes it,
never from the sequential state -- and checks that the result equals the sequential lexicographic sweeps bit for bit,
for 1, 2 and 4 columns per step, odd widths, levels narrower than a step, several 32-row blocks (whose warps skip the
steps at which none of their rows is inside the image) and 1 ... 5 sweeps.  It needs no GPU: it pins the schedule,
the GPU parity tests pin the CUDA code that implements it."""
import numpy as np
import pytest

f32 = np.float32


def lexicographic(A, B, du0, T, omega):
    """T Gauss-Seidel sweeps in raster order with the kernel's per-pixel expression (= solver.c:122-131 / 180-190 /
    237-247 after the block inversion of the first sweep): A[j, i] = (a11, a12, a22, horiz), B[j, i] = (b1, b2, vert)."""
    h, w = A.shape[:2]
    du = du0.copy()
    for _ in range(T):
        for j in range(h):
            for i in range(w):
                a11, a12, a22, hr = A[j, i]
                b1, b2, vb = B[j, i]
                old1 = du[j, i + 1] if i + 1 < w else np.zeros(2, f32)
                px, py = f32(hr * old1[0]), f32(hr * old1[1])
                if j > 0:
                    vt = B[j - 1, i, 2]
                    px, py = f32(px + f32(vt * du[j - 1, i, 0])), f32(py + f32(vt * du[j - 1, i, 1]))
                if j < h - 1:
                    px, py = f32(px + f32(vb * du[j + 1, i, 0])), f32(py + f32(vb * du[j + 1, i, 1]))
                s1, s2 = f32(px + b1), f32(py + b2)
                if i > 0:
                    hl = A[j, i - 1, 3]
                    s1, s2 = f32(f32(hl * du[j, i - 1, 0]) + s1), f32(f32(hl * du[j, i - 1, 1]) + s2)
                sx, sy = du[j, i]
                nx = f32(sx + f32(omega * f32(f32(f32(a11 * s1) + f32(a12 * s2)) - sx)))
                ny = f32(sy + f32(omega * f32(f32(f32(a12 * s1) + f32(a22 * s2)) - sy)))
                du[j, i] = (nx, ny)
    return du


def scheduled(A, B, du0, T, omega, C):
    """The same sweeps in the order and with the operand sources of k_sor_small<C>."""
    h, w = A.shape[:2]
    K = (h + 31) // 32
    rows = K * 32
    nst = (w + C - 1) // C + h - 1 + 2 * (T - 1)
    z2 = np.zeros(2, f32)
    res = np.zeros((2, T, rows, C, 2), f32)           # published results, double-buffered by step parity
    res1 = np.zeros((T, rows, 2), f32)                # (i0 - 1, j): the thread's last result
    hl = np.zeros((T, rows), f32)                     # horiz(i0 - 1, j)
    Cc = np.zeros((T, rows, C, 2), f32)               # the previous sweep at the thread's C pixels
    out = np.full((h, w, 2), np.nan, f32)

    def init(j, i):  # du before the launch; the kernel's ring holds garbage outside the row, never used; zero past it
        return du0[j, i] if 0 <= i < w and j < h else z2

    for j in range(rows):                              # sweep 0, step 0: ring slots 0 ... C - 1
        for q in range(C):
            Cc[0, j, q] = init(j, -C * j + q)
    for t in range(nst):
        cur, prev = res[t & 1], res[(t & 1) ^ 1].copy()
        for s in range(T):
            for k in range(K):
                live = (32 * k + 2 * s - 1) <= t < (32 * k + 31 + 2 * s + (w + C - 1) // C)
                for l in range(32):
                    j = 32 * k + l
                    if not live:                       # the warp only publishes zeros
                        cur[s, j] = 0
                        continue
                    i0 = C * (t - j - 2 * s)
                    jp, jn = max(j - 1, 0), min(j + 1, rows - 1)
                    U = prev[s, jp]
                    if s == 0:
                        P = np.array([init(j, i0 + C + q) for q in range(C)], f32)
                        Bl = np.array([init(jn, i0 + q) for q in range(C)], f32)
                    else:
                        P, Bl = prev[s - 1, j], prev[s - 1, jn]
                    left, hq = res1[s, j].copy(), hl[s, j]
                    nv = np.zeros((C, 2), f32)
                    for q in range(C):
                        i = i0 + q
                        act = j < h and 0 <= i < w
                        if 0 <= i < w and j < h:
                            a11, a12, a22, hr = A[j, i]
                            b1, b2, vb = B[j, i]
                            vt = B[jp, i, 2] if j > 0 else f32(0)
                        else:                          # ring garbage: any finite value, the result is discarded
                            a11 = a12 = a22 = hr = b1 = b2 = vb = vt = f32(0.5)
                        old1 = Cc[s, j, q + 1] if q < C - 1 else P[0]
                        px, py = f32(hr * old1[0]), f32(hr * old1[1])
                        if j != 0:
                            px, py = f32(px + f32(vt * U[q, 0])), f32(py + f32(vt * U[q, 1]))
                        if not j >= h - 1:
                            px, py = f32(px + f32(vb * Bl[q, 0])), f32(py + f32(vb * Bl[q, 1]))
                        s1, s2 = f32(px + b1), f32(py + b2)
                        if not (q == 0 and i0 == 0):
                            s1, s2 = f32(f32(hq * left[0]) + s1), f32(f32(hq * left[1]) + s2)
                        sx, sy = Cc[s, j, q]
                        v = np.array([f32(sx + f32(omega * f32(f32(f32(a11 * s1) + f32(a12 * s2)) - sx))),
                                      f32(sy + f32(omega * f32(f32(f32(a12 * s1) + f32(a22 * s2)) - sy)))], f32)
                        if not act:
                            v = z2.copy()
                        nv[q] = v
                        left, hq = v, hr
                        if s == T - 1 and act:
                            assert np.isnan(out[j, i, 0]), "pixel written twice"
                            out[j, i] = v
                    res1[s, j], hl[s, j] = left, hq
                    Cc[s, j] = P
                    cur[s, j] = nv
    return out


@pytest.mark.parametrize("w,h,T,C", [(13, 9, 3, 2), (12, 7, 3, 1), (9, 11, 3, 4), (3, 6, 2, 4), (2, 5, 1, 2), (17, 40, 3, 2),
                                     (21, 37, 4, 2), (10, 70, 2, 4), (7, 5, 5, 1)])
def test_small_level_schedule_is_the_lexicographic_sweep(w, h, T, C):
    rng = np.random.default_rng(w * 100 + h * 10 + T + C)
    A = rng.uniform(0.05, 1.0, (h, w, 4)).astype(f32)
    A[:, :, 1] *= f32(0.3)
    A[:, w - 1, 3] = 0            # horiz(i, j) is zero in the last column (compute_smoothness)
    B = rng.uniform(-1.0, 1.0, (h, w, 3)).astype(f32)
    B[:, :, 2] = np.abs(B[:, :, 2])
    B[h - 1, :, 2] = 0            # vert(i, j) is zero in the last row
    du0 = (rng.standard_normal((h, w, 2)) * 0.1).astype(f32)
    omega = f32(1.6)
    with np.errstate(over="ignore"):
        ref = lexicographic(A, B, du0, T, omega)
        got = scheduled(A, B, du0, T, omega, C)
    assert not np.isnan(got).any(), "a pixel was never written by the last sweep"
    assert (got.view(np.uint32) != ref.view(np.uint32)).sum() == 0


def wavefront_tickets(T, K):
    """Ticket -> (sweep t, row block k) as k_sor_wavefront decodes it: items ordered by key = 2 t + k, sweeps ascending
    inside a key (varref.cu, 'decode ticket')."""
    order = []
    for key in range(2 * (T - 1) + K):
        for t in range(T):
            k = key - 2 * t
            if 0 <= k < K:
                order.append((t, k))
    return order


@pytest.mark.parametrize("T,K", [(1, 1), (3, 1), (3, 9), (1, 68), (5, 4), (40, 5), (256, 2)])
def test_wavefront_tickets_respect_the_dependencies(T, K):
    """Persistent warps take tickets in this order and only after finishing their item, so the launch cannot deadlock
    for ANY number of CTAs iff every item an item waits for -- the block above in the same sweep, the same block and
    the block below in the previous sweep -- holds a smaller ticket."""
    order = wavefront_tickets(T, K)
    assert len(order) == T * K and len(set(order)) == T * K
    ticket = {it: n for n, it in enumerate(order)}
    for (t, k), n in ticket.items():
        for dep in ((t, k - 1), (t - 1, k), (t - 1, k + 1)):
            if dep in ticket:
                assert ticket[dep] < n, ((t, k), dep)
