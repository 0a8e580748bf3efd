"""The arithmetic claims the lock-step kernels rest on, restated in numpy / exact rationals and checked on the CPU (the kernels
themselves are compared with the oracle on the GPU; these tests pin the *reasons* they may skip work):

* K2 `axis_limit` (csrc/pvp_motion.cu): tm0 + rem*td minus 2^-11 (|tm0| + rem*td) is a lower bound of the float32 value t_max has
  after `rem` rounded additions of t_delta -- so a step whose t is below it cannot be the one that leaves the grid.
* K3 `voxel_index_screened` (csrc/pvp_image.cu): when the 20 fraction bits read from the mantissa of a*K + (2^52 - 5) are below
  0xFFFF6, the integer part read from it equals the reference's min(trunc(RN(RN(a/b) * n)), n - 1).
* K3 `sample_keys2<2>`: the same read-out from ONE fused multiply-add y = RN(s*dx + x0) per axis equals the reference's index for
  samples inside the grid whenever A*K <= 2^49 (the host's `affine_ok`) and the safe range's 32-unit margin holds.
* K3 `div_by_extent`: the five-operation FMA sequence is the correctly rounded quotient (exact rationals).
* K2 axis choice from one three-input minimum == the reference's strict-comparison chain; K2 `slab_reach`: the box-exit x index
  bounds the planes a ray visits; K2 `slab_enter`: the fast-forward to a slab's near face (three independent scalar loops) reproduces the
  reference loop's state after its k-th x step, ties and early ends included."""
from fractions import Fraction

import numpy as np

F32 = np.float32


def test_k2_axis_limit_is_a_lower_bound_of_the_iterated_sum():
    rng = np.random.default_rng(7)
    n = 20000
    tm0 = (rng.uniform(-1e3, 1e4, n) * rng.choice([1.0, 1e-3, 1e3], n)).astype(F32)
    td = np.exp(rng.uniform(np.log(1e-4), np.log(1e5), n)).astype(F32)
    rem = rng.integers(0, 2048, n)
    s = tm0.copy()
    worst = 0.0
    for k in range(2047):                       # S_k = RN(S_{k-1} + td), frozen once a sample has done its `rem` additions
        nxt = (s + td).astype(F32)
        s = np.where(k < rem, nxt, s)
    span = (rem.astype(F32) * td).astype(F32)
    closed = (tm0 + span).astype(F32)
    bound = (closed - ((np.abs(tm0) + span).astype(F32) * F32(2.0 ** -11)).astype(F32)).astype(F32)
    assert np.all(bound <= s)
    proven = (np.abs(tm0) + span) * 2.0 ** -13           # the error bound of the derivation; the kernel's slack is 4x that
    worst = float(np.max(np.abs(s.astype(np.float64) - (tm0.astype(np.float64) + rem * td.astype(np.float64))) / np.maximum(proven, 1e-30)))
    assert worst < 1.0, worst


def _ref_index(a, b, n):
    return np.minimum(np.trunc((a / b) * n).astype(np.int64), n - 1)      # process_image.cpp:94-104 (float64, RN each)


def _read_out(y):
    bits = y.view(np.uint64)
    lo20 = bits & np.uint64(0xFFFFF)
    idx = ((bits >> np.uint64(20)) & np.uint64(0xFFFFFFFF)).astype(np.int64)   # the funnel shift of the two mantissa words
    return idx, lo20 < 0xFFFF6


def test_k3_mantissa_read_out_equals_the_reference_index():
    rng = np.random.default_rng(11)
    checked = 0
    for trial in range(60):
        b = float(np.exp(rng.uniform(-8, 12)))
        n = int(rng.choice([1, 2, 7, 64, 400, 512, 4096, 1 << 20, 1 << 24]))
        K = (n / b) * 1048576.0
        a = np.concatenate([rng.uniform(0, b, 200000), np.array([0.0, b])])
        k = rng.integers(0, n + 1, 20000).astype(np.float64)                    # voxel boundaries and their neighbourhoods
        edge = (k / n) * b
        a = np.concatenate([a, edge, np.nextafter(edge, 0), np.nextafter(edge, 2 * b), edge * (1 + 3e-7), edge * (1 - 3e-7)])
        a = a[(a >= 0) & (a <= b)]
        y = a * K + (2.0 ** 52 - 5.0)
        idx, safe = _read_out(y)
        want = _ref_index(a, b, n)
        assert np.array_equal(np.minimum(idx[safe], n - 1), want[safe]), (b, n)
        assert safe.mean() > 0.5 or n >= 1 << 20 or b < 1e-3
        checked += int(safe.sum())
    assert checked > 5_000_000


def test_k3_affine_coordinate_equals_the_reference_index_inside_the_safe_range():
    rng = np.random.default_rng(13)
    checked = flagged = 0
    for trial in range(600):
        n = int(rng.choice([16, 400, 512, 2048]))
        b = float(rng.uniform(0.5, 4000.0))
        lo = float(rng.uniform(-3000.0, 3000.0))
        hi = lo + b
        bb = hi - lo                                    # :94-96 divides by RN(max - min)
        K = (n / bb) * 1048576.0
        o = float(rng.uniform(lo - 3 * b, hi + 3 * b))
        w = float(rng.uniform(-1, 1)) or 0.5
        step = float(rng.uniform(0.01, 2.0) * b / n)
        max_distance = 8 * b
        A = abs(o) + max_distance + abs(lo) + abs(hi) + bb
        if not A * K <= 2.0 ** 49:                      # make_fastp: affine_ok
            continue
        dx = (step * w) * K                             # affine_of()
        x0 = (o - lo) * K + (2.0 ** 52 - 5.0)
        margin = 2.0 ** -47 * (abs(o) + max_distance * abs(w) + abs(lo) + abs(hi)) + 32.0 / K    # safe_range: 2e + 32 units
        for s in rng.integers(0, int(max_distance / step) + 1, 250):
            s = int(s)
            q = o + (s * step) * w                      # :344-347, RN each
            if not (lo + margin <= q <= hi - margin):
                continue
            y = float(Fraction(s) * Fraction(dx) + Fraction(x0))   # one fused multiply-add: exact product and sum, rounded once
            idx, safe = _read_out(np.array([y]))
            if not safe[0]:
                flagged += 1                            # the kernel takes the exact sequence there
                continue
            want = _ref_index(np.array([q - lo]), bb, n)[0]
            assert idx[0] == want, (trial, s, n, b, o, w)
            checked += 1
    assert checked > 10000 and flagged < checked // 50


def test_k3_affine_coordinate_near_voxel_faces_at_the_largest_allowed_magnitudes():
    """The hard corner of the claim above: sample positions steered to within a few 2^-20 voxel units of a voxel face, with
    A*K between 2^46 and 2^49 (far observers, fine grids), where the affine value and the reference's own rounding differ most."""
    rng = np.random.default_rng(17)
    checked = flagged = 0
    while checked + flagged < 6000:
        n = int(rng.choice([512, 4096, 1 << 16, 1 << 20]))
        b = float(rng.uniform(1.0, 1000.0))
        lo = float(rng.uniform(-1000.0, 1000.0))
        hi = lo + b
        bb = hi - lo
        K = (n / bb) * 1048576.0
        target = 2.0 ** rng.uniform(46, 49)             # wanted A*K
        far = target / K                                # ~ |o| + max_distance
        w = float(rng.choice([-1, 1]) * rng.uniform(0.3, 1.0))
        max_distance = 0.5 * far
        step = max_distance / 4096
        s = int(rng.integers(1, 4096))
        k = int(rng.integers(1, n - 1))
        delta = float(rng.integers(-12, 13)) + float(rng.uniform(-0.5, 0.5))
        o = lo + (k + delta * 2.0 ** -20) * bb / n - (s * step) * w     # steer sample s to voxel face k (+ delta units)
        A = abs(o) + max_distance + abs(lo) + abs(hi) + bb
        if not (A * K <= 2.0 ** 49):
            continue
        q = o + (s * step) * w
        margin = 2.0 ** -47 * (abs(o) + max_distance * abs(w) + abs(lo) + abs(hi)) + 32.0 / K
        if not (lo + margin <= q <= hi - margin):
            continue
        dx = (step * w) * K
        x0 = (o - lo) * K + (2.0 ** 52 - 5.0)
        y = float(Fraction(s) * Fraction(dx) + Fraction(x0))
        idx, safe = _read_out(np.array([y]))
        if not safe[0]:
            flagged += 1
            continue
        assert idx[0] == _ref_index(np.array([q - lo]), bb, n)[0], (n, b, lo, o, w, s, k, delta)
        checked += 1
    assert checked > 2000 and flagged > 500             # both outcomes exercised: the window really sits on the faces


def test_k2_slab_reach_never_skips_a_ray_that_enters_the_slab():
    """K2 `slab_reach` (csrc/pvp_motion.cu): a rank skips a ray when the x index of its box-exit point, two voxels of margin
    added, stays short of the rank's slab -- x_exit = (cam_x + t_end * dir_x - gmin_x) / vs in float32, skip iff
    x_exit + 2 < slab_lo (walking +x) or x_exit - 2 >= slab_hi (walking -x).  Restated here with the reference's own slab test
    for t_end (ray_voxel.cpp:296-317, float32) and checked against the voxel lists of the oracle: the furthest x index a ray
    really visits is never beyond that estimate, for cameras inside and outside, grazing and axis-parallel rays."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import Oracle
    o = Oracle()
    rng = np.random.default_rng(23)
    slack = []
    n_rays = 0
    for k in range(14000):
        N = int(rng.integers(2, 96))
        vs = F32(rng.choice([0.37, 1.0, 2.0, 6.0]) if k % 2 else rng.uniform(0.05, 9.0))
        center = rng.uniform(-300, 300, 3).astype(F32)
        half = F32(0.5) * (F32(N) * vs)
        gmin, gmax = (center - half).astype(F32), (center + half).astype(F32)           # make_gridp
        cam = (center + rng.uniform(-2.0, 2.0, 3) * half).astype(F32)
        d = ((center + rng.uniform(-1.1, 1.1, 3) * half) - cam).astype(F32) if k % 3 else rng.normal(size=3).astype(F32)
        if k % 5 == 0:
            d[rng.integers(1, 3)] = F32(rng.choice([0.0, -0.0, 1e-13]))                  # parallel to y or z
        if k % 11 == 0:
            d[0] *= F32(1e-4)                                                            # nearly parallel to the x-planes: grazing in x
        nrm = np.sqrt(np.sum(d.astype(np.float64) ** 2))
        if nrm == 0:
            continue
        d = (d / F32(nrm)).astype(F32)
        vox = o.cast_ray(cam, d, N, vs, center)
        if len(vox) == 0:
            continue
        # t_end: the reference's slab test, float32, fminf / fmaxf
        t_min, t_max, hit = F32(-np.inf), F32(np.inf), True
        for a in range(3):
            if abs(d[a]) < 1e-12:
                if cam[a] < gmin[a] or cam[a] > gmax[a]:
                    hit = False
                continue
            t1, t2 = F32((gmin[a] - cam[a]) / d[a]), F32((gmax[a] - cam[a]) / d[a])
            t_min, t_max = max(t_min, min(t1, t2)), min(t_max, max(t1, t2))
        assert hit and t_min <= t_max
        x_exit = F32(F32(F32(cam[0] + F32(t_max * d[0])) - gmin[0]) / vs)
        ix = vox[:, 0]
        if d[0] >= 0:           # sx = +1 (ray_voxel.cpp:337: dir >= 0, incl. -0.0)
            assert not (F32(x_exit + F32(2.0)) < F32(ix.max())), (k, x_exit, ix.max())      # a slab starting at the last visited plane is reached
            slack.append(float(ix.max()) - float(x_exit))
        else:
            assert not (F32(x_exit - F32(2.0)) >= F32(ix.min() + 1)), (k, x_exit, ix.min())
            slack.append(float(x_exit) - float(ix.min() + 1))
        n_rays += 1
    assert n_rays > 5000
    assert max(slack) < 1.0, max(slack)     # the estimate is within one voxel; the kernel allows two


def test_k2_axis_choice_by_min3_equals_the_reference_chain():
    """K2's DDA step picks the axis from one three-input minimum: z iff tmz == m, else y iff tmy == m, else x (m = min of the
    three).  The reference's chain is `x if tx < ty && tx < tz; else y if ty < tz; else z` (ray_voxel.cpp:375-387: strict
    comparisons, ties go to the later axis).  Equal for every NaN-free triple, including ties, both zeros and infinities (the
    host keeps NaN out: lean_geometry_ok)."""
    vals = np.array([0.0, -0.0, 1.0, 1.0 + 2.0 ** -23, 2.0, -1.0, np.inf, 2.0 ** -149, 3.5, 1e30, -np.inf], F32)
    tx, ty, tz = (a.ravel() for a in np.meshgrid(vals, vals, vals, indexing="ij"))
    ref = np.where((tx < ty) & (tx < tz), 0, np.where(ty < tz, 1, 2))
    m = np.minimum(np.minimum(tx, ty), tz)
    cz = tz == m
    cy = ~cz & (ty == m)
    ours = np.where(cz, 2, np.where(cy, 1, 0))
    assert np.array_equal(ref, ours) and len(ref) == len(vals) ** 3
    rng = np.random.default_rng(3)
    t = rng.integers(0, 6, (200000, 3)).astype(F32) * F32(0.25)          # many ties
    ref = np.where((t[:, 0] < t[:, 1]) & (t[:, 0] < t[:, 2]), 0, np.where(t[:, 1] < t[:, 2], 1, 2))
    m = t.min(1)
    ours = np.where(t[:, 2] == m, 2, np.where(t[:, 1] == m, 1, 0))
    assert np.array_equal(ref, ours)


def test_k3_division_by_extent_is_the_correctly_rounded_quotient():
    """K3 `div_by_extent` (csrc/pvp_image.cu): q = a*y; r = fma(-q, b, a); q = fma(r, y, q); r = fma(-q, b, a); q = fma(r, y, q)
    with y = RN(1/b) equals RN(a / b) -- the reference's `/` (process_image.cpp:94-96) -- for every divisor the host admits
    (`make_fastp`: significand not all ones).  Restated with exact rationals (a fused multiply-add is one rounding of the
    exact value); the GPU self-test runs the real instruction sequence on 4e7 numerators per extent."""
    def rn(x):
        return float(x)                      # Fraction -> float: one correctly rounded conversion

    def fma(a, b, c):
        return rn(Fraction(a) * Fraction(b) + Fraction(c))

    rng = np.random.default_rng(41)
    extents = [20.0, 17.0, 3.0, 1.0 / 3.0, 1e-3, 6e12, 2.0 ** 40 + 1.0, 1.9999999999999996, 1.0000000000000002, 0.7, 123456.789, 5e-324 * 2 ** 600,
               float(np.nextafter(4.0, 0.0) / 1.0), 400.0 * 0.1, np.pi, 2.0 / 3.0]
    checked = 0
    for b in extents:
        m, _ = np.frexp(b)
        if m == 1.0 - 2.0 ** -53:            # make_fastp refuses this divisor (the case Markstein's theorem excludes)
            continue
        y = 1.0 / b
        n = 512
        a_list = list(rng.uniform(0.0, b, 300)) + [b * k / n for k in range(0, n + 1, 37)]
        for k in (1, 2, 255, 256, 511):      # the neighbourhood of voxel boundaries, where a wrong last bit would change an index
            base = b * k / n
            a_list += [float(np.nextafter(base, 0.0)), base, float(np.nextafter(base, np.inf))]
        a_list += [0.0, b, float(np.nextafter(b, 0.0)), 5e-324, b * 2.0 ** -60]
        for a in a_list:
            q = a * y
            r = fma(-q, b, a)
            q = fma(r, y, q)
            r = fma(-q, b, a)
            q = fma(r, y, q)
            assert q == a / b, (a, b, q, a / b)
            checked += 1
    assert checked > 4000


def test_k3_safe_range_samples_pass_the_reference_bounds_test():
    """K3 `safe_range` (csrc/pvp_image.cu): (1) the computed sample position q(s) = RN(o + RN(RN(s*step) * w)) of
    process_image.cpp:344-347 lies within 2^-48 (|o| + d_max |w|) of the exact, linear x(s) = o + s*step*w (exact rationals);
    (2) therefore, when the first and last sample of a span keep the margin 2^-47 (|o| + d_max |w| + |ext_lo| + |ext_hi|) to all
    six faces, EVERY computed position between them passes the inclusive bounds test of :85 -- which is what lets the march skip
    the test there.  (2) is checked by brute force on spans the restated routine accepts, incl. rays that graze a face."""
    rng = np.random.default_rng(97)
    # (1) the error bound
    for _ in range(3000):
        scale = float(rng.choice([1.0, 1e3, 6e12]))
        o = float(rng.uniform(-scale, scale))
        w = float(rng.uniform(-1, 1))
        step = float(rng.uniform(1e-3, 10.0) * scale / 1e3)
        s = int(rng.integers(0, 20001))
        d = float(s) * step
        q = o + d * w
        exact = Fraction(o) + Fraction(s) * Fraction(step) * Fraction(w)
        bound = Fraction(2) ** -48 * (abs(Fraction(o)) + abs(Fraction(d) * Fraction(w)))
        assert abs(Fraction(q) - exact) <= bound, (o, w, step, s)
    # (2) spans accepted by the routine contain only samples that pass the reference's test
    accepted = grazing = 0
    for case in range(400):
        scale = float(rng.choice([1.0, 50.0, 6e12]))
        ext = np.sort(rng.uniform(-scale, scale, (3, 2)), axis=1)
        o = rng.uniform(-2 * scale, 2 * scale, 3)
        target = ext[:, 0] + rng.uniform(0.02, 0.98, 3) * (ext[:, 1] - ext[:, 0])
        w = target - o
        if case % 4 == 0:                       # run along a face: one coordinate sits (almost) on ext_lo all the way
            k = int(rng.integers(0, 3))
            o[k] = ext[k, 0] + (ext[k, 1] - ext[k, 0]) * float(rng.choice([0.0, 1e-18, 1e-13]))
            w[k] = 0.0
            grazing += 1
        w = w / np.sqrt(np.sum(w * w))
        n_steps = int(rng.choice([200, 2000, 20000]))
        step = 4.0 * scale * 2 / n_steps
        s_all = np.arange(0, n_steps + 1, dtype=np.float64)
        d = s_all * step
        q = o[None, :] + d[:, None] * w[None, :]                                             # :344-347, float64, RN each
        ok = np.all((q >= ext[None, :, 0]) & (q <= ext[None, :, 1]), axis=1)                 # :85 inclusive
        inside_idx = np.flatnonzero(ok)
        if len(inside_idx) < 8:
            continue
        lo, hi = int(inside_idx[0]), int(inside_idx[-1])
        dmax = abs(hi * step) + abs(lo * step)
        e = 2.0 ** -47 * (np.abs(o) + dmax * np.abs(w) + np.abs(ext[:, 0]) + np.abs(ext[:, 1]))

        def inside(s):
            qq = o + (float(s) * step) * w
            return bool(np.all((ext[:, 0] + e <= qq) & (qq <= ext[:, 1] - e)))
        a, b = lo, hi
        for _ in range(3):
            if a <= hi and not inside(a):
                a += 1
        for _ in range(3):
            if b >= lo and not inside(b):
                b -= 1
        if a <= b and inside(a) and inside(b):
            accepted += 1
            assert ok[a:b + 1].all(), (case, a, b)
    assert accepted > 150 and grazing > 50


def test_k2_slab_fast_forward_is_the_reference_walk():
    """`slab_enter` (csrc/pvp_motion.cu): the state of the reference's DDA loop (ray_voxel.cpp:365-391) right after its k-th x step --
    voxel, the three t_max, t_current, emitted voxels -- equals the three independent loops of the fast-forward: X = S_x[k-1]; every y step
    with S_y[j] <= X and every z step with S_z[j] <= X precede it (ties go to the later axis); the ray dies first if X > t_end or if y / z
    leave [0, N).  float32 throughout, incl. exact ties (equal t_max on two axes, t_delta that are multiples of each other) and parallel
    axes (t_max = inf)."""
    rng = np.random.default_rng(21)
    N = 40
    checked = died = 0
    for trial in range(4000):
        tie = trial % 5 == 0
        td = (np.array([1.0, 2.0, 0.5], np.float32) * F32(rng.choice([0.25, 1.0, 3.0])) if tie
              else np.exp(rng.uniform(np.log(0.05), np.log(20.0), 3)).astype(F32))
        tm = (td * rng.integers(1, 4, 3).astype(F32)).astype(F32) if tie else (rng.uniform(0.0, 1.0, 3).astype(F32) * td).astype(F32)
        par = rng.integers(0, 12)
        if par in (1, 2):                                     # y or z parallel to the ray: never steps (safe_div -> inf)
            tm[par], td[par] = np.inf, np.inf
        idx = rng.integers(0, N, 3)
        step = rng.choice([-1, 1], 3)
        t_end = F32(rng.uniform(5.0, 200.0))
        k = int(rng.integers(1, N))                          # the slab's near face is k x-steps away
        if not (0 <= idx[0] + step[0] * k < N):
            continue
        # ---- the reference loop, run until its k-th x step (or until it ends)
        t_cur, m, i, xs, emitted, alive = F32(0.0), tm.copy(), idx.copy(), 0, 0, True
        while True:
            if not (t_cur <= t_end):
                alive = False
                break
            emitted += 1
            if m[0] < m[1] and m[0] < m[2]:
                a = 0
            elif m[1] < m[2]:
                a = 1
            else:
                a = 2
            i[a] += step[a]
            t_cur = m[a]
            m[a] = F32(m[a] + td[a])
            if not (0 <= i[a] < N):
                alive = False
                break
            if a == 0:
                xs += 1
                if xs == k:
                    break
        # ---- the fast-forward
        X = tm[0]
        for _ in range(1, k):
            X = F32(X + td[0])
        ff_alive = bool(X <= t_end)
        cnt, val = [0, 0], [tm[1], tm[2]]
        if ff_alive:
            for ax in (1, 2):
                rem = N - 1 - idx[ax] if step[ax] > 0 else idx[ax]
                while val[ax - 1] <= X:
                    cnt[ax - 1] += 1
                    if cnt[ax - 1] > rem:
                        ff_alive = False
                        break
                    val[ax - 1] = F32(val[ax - 1] + td[ax])
                if not ff_alive:
                    break
        # the loop only notices "X > t_end" at the NEXT iteration's check: the k-th x step itself is taken, the voxel after it is not emitted
        ref_enters = alive and bool(t_cur <= t_end)
        assert ref_enters == ff_alive, (trial, alive, ff_alive)
        if ff_alive:
            checked += 1
            assert emitted == k + cnt[0] + cnt[1]
            assert i[0] == idx[0] + step[0] * k and i[1] == idx[1] + step[1] * cnt[0] and i[2] == idx[2] + step[2] * cnt[1]
            assert t_cur == X and m[0] == F32(X + td[0]) and m[1] == val[0] and m[2] == val[1]
        else:
            died += 1
    assert checked > 500 and died > 200, (checked, died)


def _lane_masks(lane, cap):
    """csrc/pvp_motion.cu lane_masks<CAP> / csrc/pvp_image.cu scan_masks_cap<CAP>, restated."""
    def win(d):
        return (((1 << d) - 1) << (lane + 1 - d)) if lane + 1 >= d else ((2 << lane) - 1)
    m = {1: 1 << lane, 2: win(2), 4: win(4), 8: win(8), 16: win(16), "next": 1 if lane == 31 else 1 << (lane + 1)}
    if cap < 32:
        g = lane % cap
        for d in (1, 2, 4, 8, 16):
            if d > g:
                m[d] |= 1
        if g == cap - 1:
            m["next"] |= 1
    return m


def test_lockstep_merge_machinery_conserves_every_voxel_sum():
    """lean4 / the N-pixel K3 restated on the CPU: per lane a bottom-up chain (ray j joins ray j-1 where their voxel keys are equal,
    else issues its own RED), across lanes a segmented scan over run heads of ray 0's key whose runs are CUT at every CAP-th lane
    (the cut is encoded in the lane masks only: bit 0 of the head ballot is always set), one RED per run from the run's last lane.
    Claim: for ANY keys -- parked rays (NONE) included, equal keys that are not adjacent included -- the REDs issued add up, per voxel,
    to exactly the sum of the values of the rays that sit in it; and no RED goes to NONE."""
    rng = np.random.default_rng(11)
    NONE = 0xFFFFFFFF
    for cap, nr in ((32, 2), (8, 4), (4, 3), (16, 4), (8, 1)):
        levels = [d for d in (1, 2, 4, 8, 16) if d < cap]
        masks = [_lane_masks(l, cap) for l in range(32)]
        for trial in range(300):
            # keys with long runs, short runs, repeats and parked rays
            n_vox = int(rng.integers(1, 12))
            base = np.sort(rng.integers(0, n_vox, 32))
            keys = np.empty((32, nr), np.int64)
            for j in range(nr):
                keys[:, j] = base + (rng.random(32) < 0.15 * j)          # vertical neighbours mostly share the voxel
            if trial % 3 == 0:
                rng.shuffle(keys)                                        # runs in any order, equal keys not adjacent
            keys[rng.random((32, nr)) < 0.1] = NONE
            vals = rng.integers(1, 256, (32, nr)).astype(np.int64)
            vals[keys == NONE] = 0                                       # a parked ray carries value 0
            want = {}
            for k, v in zip(keys.ravel(), vals.ravel()):
                if k != NONE:
                    want[int(k)] = want.get(int(k), 0) + int(v)
            got = {}

            def red(k, v):
                assert k != NONE
                got[int(k)] = got.get(int(k), 0) + int(v)

            s = vals.copy()
            for l in range(32):                                          # the lane's chain
                for j in range(nr - 1, 0, -1):
                    same = keys[l, j] == keys[l, j - 1]
                    if not same and keys[l, j] != NONE:
                        red(keys[l, j], s[l, j])
                    if same:
                        s[l, j - 1] += s[l, j]
            k0 = keys[:, 0]
            heads = 0
            for l in range(32):                                          # run_head: lane 0, or a key that differs from the lane below
                if l == 0 or k0[l] != k0[l - 1]:
                    heads |= 1 << l
            assert heads & 1
            run = s[:, 0].copy()
            for d in levels:                                             # segmented scan, all lanes at once per level
                prev = run.copy()
                for l in range(32):
                    if not (heads & masks[l][d]) and l >= d:
                        run[l] = prev[l] + prev[l - d]
                    else:
                        assert (heads & masks[l][d]) or l < d            # (a window that reaches below lane 0 always contains bit 0)
            for l in range(32):
                if (heads & masks[l]["next"]) and k0[l] != NONE:
                    red(k0[l], run[l])
            assert got == want, (cap, nr, trial)


def test_fp32_hot_cell_error_grows_like_the_square_root_of_the_adds():
    """The tolerance of tests/test_gpu_fullsize_parity.py for cells above 2^24 (hot_tol = 1e-5 per 2^20 rays, growing with the square root
    of the ray count) restated: the camera's own voxel receives one RED per warp chunk (~32 rays of <= 255 each), every RED into an fp32
    accumulator above 2^24 rounds by up to half an ulp, and the roundings add up like a random walk.  Sequential float32 sums
    (np.cumsum) of K = rays / 32 integer run sums against the exact integer sum, 40 orders each: the error stays below hot_tol, and it
    is not far below it -- i.e. the tolerance is neither loose by an order of magnitude nor reachable at a fixed 1e-5 for 4K frames."""
    rng = np.random.default_rng(5)
    for n_rays in (1 << 21, 1 << 23):                       # a 1080p pair, a 4K pair
        tol = 1e-5 * max(1.0, (n_rays / 2.0 ** 20) ** 0.5)
        k = n_rays // 32
        worst = 0.0
        for _ in range(40):
            runs = rng.integers(1, 256, (k, 32)).sum(axis=1)     # run sums of 32 rays, exact integers < 2^24
            exact = int(runs.sum())
            got = float(np.cumsum(runs.astype(np.float32), dtype=np.float32)[-1])
            worst = max(worst, abs(got - exact) / exact)
        assert exact > 2 ** 24 and worst <= tol, (n_rays, worst, tol)
        assert worst >= tol / 30, (n_rays, worst, tol)           # the same order of magnitude: 2.5e-6 ... at 1080p
