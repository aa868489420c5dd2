"""The unit list of the radix-4 (latency-mode) G1-NTT kernel, k_fk20_g1_ntts_r4 in csrc/kzg_kernels.cu, replayed on the CPU with
elements of Fr standing in for the points (both are modules over Fr, the schedule only adds and multiplies by roots of
unity): the seven super-phases must give exactly what the fourteen radix-2 phases of k_fk20_g1_ntts give -- which the
consensus vectors pin -- and both must equal the definition: inverse transform of the bit-reversed input, first 64
coefficients kept, forward transform of (h || 0), output bit-reversed (reference: fk20/prover.rs:199-222 over
polynomial/src/domain.rs:149-194).  The index arithmetic below is a transcription of r4_mul_unit / r4_combine_unit."""
import random

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
W = pow(7, (R - 1) // 128, R)          # omega_128 (7 generates the multiplicative group; only "a primitive 128th root" matters here)
TW = [pow(W, e, R) for e in range(128)]


def rev7(x):
    return int(format(x, "07b")[::-1], 2)


def radix2_phases(p):
    """k_fk20_g1_ntts: g1_ntt_butterfly for ph = 0..13"""
    p = list(p)
    for ph in range(14):
        mode, st = (1, 13 - ph) if ph >= 7 else (0, ph)
        ln = 1 << st
        for t in range(64):
            pos = t & (ln - 1)
            i = ((t >> st) << (st + 1)) + pos
            j = i + ln
            e = pos << (6 - st)
            if mode == 0:
                u, v = p[i], p[j] * TW[(128 - e) & 127] % R
                p[i] = (u + v) % R
                if st != 6:
                    p[j] = (u - v) % R
            elif st == 6:
                p[j] = p[i] * TW[e] % R
            else:
                u, v = p[i], p[j]
                p[i] = (u + v) % R
                p[j] = (u - v) * TW[e] % R
    return p


def radix4_superphases(p):
    p = list(p)
    for sp in range(7):
        tmp = [None] * 160
        nmul = 128 if sp == 3 else 160
        for u in range(nmul):                       # r4_mul_unit
            if sp == 3:
                t = u >> 1
                tmp[u] = p[t] * TW[t] % R if u & 1 else p[t + 64] * TW[(128 - t) & 127] % R
                continue
            fwd = sp > 3
            s = 2 * (6 - sp) if fwd else 2 * sp
            ln = 1 << s
            q, which = divmod(u, 5)
            pos = q & (ln - 1)
            base = ((q >> s) << (s + 2)) + pos
            ea, eb = pos << (6 - s), pos << (5 - s)
            e = [ea, eb, (eb + 32) if fwd else (ea + eb), (ea + eb) if fwd else (eb + 32), ea + eb + 32][which]
            if not fwd:
                src = 1 if which == 0 else 2 if which in (1, 3) else 3
                tmp[u] = p[base + src * ln] * TW[(128 - e) & 127] % R
            else:
                a = 1 if which in (2, 4) else 0
                c = 1 if which == 0 else a + 2
                x = p[base + a * ln] - p[base + c * ln]
                if which == 0:
                    x += p[base + 2 * ln] - p[base + 3 * ln]
                tmp[u] = x * TW[e & 127] % R
        for c in range(192 - nmul):                 # r4_combine_unit
            if sp == 3:
                p[c], p[c + 64] = (p[c] + tmp[2 * c]) % R, (p[c + 64] + tmp[2 * c + 1]) % R
                continue
            fwd = sp > 3
            s = 2 * (6 - sp) if fwd else 2 * sp
            ln = 1 << s
            pos = c & (ln - 1)
            base = ((c >> s) << (s + 2)) + pos
            o = [base, base + ln, base + 2 * ln, base + 3 * ln]
            t = tmp[5 * c:5 * c + 5]
            if not fwd:
                a, bm = p[o[0]] + t[0], p[o[0]] - t[0]
                cc, d = t[1] + t[2], t[3] - t[4]
                p[o[0]], p[o[2]], p[o[1]], p[o[3]] = (a + cc) % R, (a - cc) % R, (bm + d) % R, (bm - d) % R
            else:
                y0 = p[o[0]] + p[o[1]] + p[o[2]] + p[o[3]]
                p[o[0]], p[o[1]], p[o[2]], p[o[3]] = y0 % R, t[0], (t[1] + t[2]) % R, (t[3] - t[4]) % R
    return p


def by_definition(m):
    """m[j] = the MSM output for frequency j (K4 stores it at rev7(j)); returns the 128 proofs in the kernel's output order, times 128
    (the 1/128 of the inverse transform is folded into the scalars by K2)"""
    h = [sum(m[j] * pow(W, (128 - (i * j) % 128) % 128, R) for j in range(128)) % R for i in range(64)]
    full = [sum(h[i] * pow(W, (i * k) % 128, R) for i in range(64)) % R for k in range(128)]
    return [full[rev7(k)] for k in range(128)]


def test_radix4_superphases_equal_radix2_phases_and_the_definition():
    rng = random.Random(4)
    for trial in range(3):
        m = [rng.randrange(R) for _ in range(128)]
        if trial == 1:
            m = [0] * 128
            m[5] = 1
        stored = [0] * 128
        for j in range(128):
            stored[rev7(j)] = m[j]
        want = by_definition(m)
        r2 = radix2_phases(stored)
        r4 = radix4_superphases(stored)
        assert r2 == want, "radix-2 phase list differs from the definition"
        assert r4 == want, "radix-4 super-phases differ from the definition"


def test_radix4_exponents_stay_inside_the_twiddle_table():
    """every twiddle index the kernel forms is in [0, 128), and the special rows 0 (copy) and 64 (negation) are the only ones
    without an op list"""
    seen = set()
    for sp in (0, 1, 2, 4, 5, 6):
        fwd = sp > 3
        s = 2 * (6 - sp) if fwd else 2 * sp
        for q in range(32):
            pos = q & ((1 << s) - 1)
            ea, eb = pos << (6 - s), pos << (5 - s)
            for e in (ea, eb, ea + eb, eb + 32, ea + eb + 32):
                assert 0 <= e < 128
                seen.add(e if fwd else (128 - e) & 127)
    assert 64 not in seen
