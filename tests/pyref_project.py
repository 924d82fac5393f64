"""A second, independent restatement (pure Python, numpy float32 scalars) of the window searches, used on small cases to
cross-check the C++ oracle: src/Frame.cc:341-356,445-510 and src/ORBmatcher.cc:45-129, 408-523, 1331-1463 of the reference.
Test infrastructure only."""
import math

import numpy as np

F = np.float32
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30
_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def dist(a, b):
    return int(_POP[np.bitwise_xor(a, b)].sum())


def cround(v):
    return int(math.floor(float(v) + 0.5)) if v >= 0 else -int(math.floor(-float(v) + 0.5))


class PyFrame:
    def __init__(self, fa, scale, blocked=None, cols=64, rows=48):
        self.__dict__.update(fa)
        self.sf, self.blocked, self.cols, self.rows = scale, blocked, cols, rows
        self.minx, self.miny, self.maxx, self.maxy = (F(v) for v in fa["bounds"])
        self.invw = F(cols) / F(self.maxx - self.minx)
        self.invh = F(rows) / F(self.maxy - self.miny)
        self.n = len(self.x)
        self.grid = [[[] for _ in range(rows)] for _ in range(cols)]
        for i in range(self.n):
            px = cround(F(F(self.x[i] - self.minx) * self.invw))
            py = cround(F(F(self.y[i] - self.miny) * self.invh))
            if 0 <= px < cols and 0 <= py < rows:
                self.grid[px][py].append(i)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = F(x), F(y), F(r)
        out = []
        a = max(0, math.floor(F(F(F(x - self.minx) - r) * self.invw)))
        if a >= self.cols:
            return out
        b = min(self.cols - 1, math.ceil(F(F(F(x - self.minx) + r) * self.invw)))
        if b < 0:
            return out
        c = max(0, math.floor(F(F(F(y - self.miny) - r) * self.invh)))
        if c >= self.rows:
            return out
        d = min(self.rows - 1, math.ceil(F(F(F(y - self.miny) + r) * self.invh)))
        if d < 0:
            return out
        check = min_level > 0 or max_level >= 0
        for ix in range(a, b + 1):
            for iy in range(c, d + 1):
                for i in self.grid[ix][iy]:
                    if check:
                        if self.octave[i] < min_level:
                            continue
                        if max_level >= 0 and self.octave[i] > max_level:
                            continue
                    if abs(F(self.x[i] - x)) < r and abs(F(self.y[i] - y)) < r:
                        out.append(i)
        return out


def three_maxima(sizes):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if F(m2) < F(0.1) * F(m1):
        i2 = i3 = -1
    elif F(m3) < F(0.1) * F(m1):
        i3 = -1
    return i1, i2, i3


def rot_bin(a1, a2):
    rot = F(F(a1) - F(a2))
    if rot < 0:
        rot = F(rot + F(360.0))
    b = cround(F(rot * F(1.0 / HISTO_LENGTH)))
    return 0 if b == HISTO_LENGTH else b


def search_projection_map(fr, mp, th, nnratio):
    owner = [-1] * fr.n
    n = 0
    for i in range(len(mp["in_view"])):
        if not mp["in_view"][i]:
            continue
        lvl = int(mp["level"][i])
        r = F(2.5) if float(mp["view_cos"][i]) > 0.998 else F(4.0)
        if float(th) != 1.0:
            r = F(r * F(th))
        rad = F(r * fr.sf[lvl])
        cand = fr.features_in_area(mp["proj_x"][i], mp["proj_y"][i], rad, lvl - 1, lvl)
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for idx in cand:
            has_obs = bool(mp["claims"][owner[idx]]) if owner[idx] >= 0 else bool(fr.blocked is not None and fr.blocked[idx])
            if has_obs:
                continue
            if fr.uright is not None and fr.uright[idx] > 0:
                if abs(F(mp["proj_xr"][i] - fr.uright[idx])) > rad:
                    continue
            d = dist(mp["desc"][i], fr.desc[idx])
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, int(fr.octave[idx]), idx
            elif d < bd2:
                bl2, bd2 = int(fr.octave[idx]), d
        if bd <= TH_HIGH:
            if bl == bl2 and F(bd) > F(F(nnratio) * F(bd2)):
                continue
            owner[bi] = i
            n += 1
    return n, np.array(owner, np.int32)


def _rt(T, p):
    return [F(F(F(F(T[r, 0] * p[0]) + F(T[r, 1] * p[1])) + F(T[r, 2] * p[2])) + T[r, 3]) for r in range(3)]


def search_projection_frame(cf, lf, mbf, mb, th, mono, check_ori):
    Tcw, Tlw = lf["Tcw"], lf["Tlw"]
    owner = [-1] * cf.n
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    twc = [F(-((float(Tcw[0, r]) * float(Tcw[0, 3]) + float(Tcw[1, r]) * float(Tcw[1, 3])) + float(Tcw[2, r]) * float(Tcw[2, 3]))) for r in range(3)]
    tlc = _rt(Tlw, twc)
    fwd = bool(tlc[2] > F(mb)) and not mono
    bwd = bool(F(-tlc[2]) > F(mb)) and not mono
    fx, fy, cx, cy = F(lf["fx"]), F(lf["fy"]), F(lf["cx"]), F(lf["cy"])
    for i in range(len(lf["has_point"])):
        if not lf["has_point"][i]:
            continue
        c = _rt(Tcw, lf["world"][i])
        with np.errstate(divide="ignore"):
            invz = F(np.float64(1.0) / np.float64(c[2]))
        if invz < 0:
            continue
        u = F(F(F(fx * c[0]) * invz) + cx)
        v = F(F(F(fy * c[1]) * invz) + cy)
        if u < cf.minx or u > cf.maxx or v < cf.miny or v > cf.maxy:
            continue
        o = int(lf["octave"][i])
        rad = F(F(th) * cf.sf[o])
        if fwd:
            cand = cf.features_in_area(u, v, rad, o, -1)
        elif bwd:
            cand = cf.features_in_area(u, v, rad, 0, o)
        else:
            cand = cf.features_in_area(u, v, rad, o - 1, o + 1)
        bd, bi = 256, -1
        for i2 in cand:
            has_obs = bool(lf["claims"][owner[i2]]) if owner[i2] >= 0 else bool(cf.blocked is not None and cf.blocked[i2])
            if has_obs:
                continue
            if cf.uright is not None and cf.uright[i2] > 0:
                ur = F(u - F(F(mbf) * invz))
                if abs(F(ur - cf.uright[i2])) > rad:
                    continue
            d = dist(lf["desc"][i], cf.desc[i2])
            if d < bd:
                bd, bi = d, i2
        if bd <= TH_HIGH:
            owner[bi] = i
            n += 1
            if check_ori:
                hist[rot_bin(lf["angle"][i], cf.angle[bi])].append(bi)
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in hist[b]:
                    owner[j] = -2
                    n -= 1
    return n, np.array(owner, np.int32)


def search_initialization(f2, f1, window, nnratio, check_ori):
    n1 = len(f1["octave1"])
    m12 = [-1] * n1
    m21 = [-1] * f2.n
    md = [2 ** 31 - 1] * f2.n
    prev = f1["prev"].copy()
    hist = [[] for _ in range(HISTO_LENGTH)]
    n = 0
    for i1 in range(n1):
        l1 = int(f1["octave1"][i1])
        if l1 > 0:
            continue
        cand = f2.features_in_area(prev[i1, 0], prev[i1, 1], F(window), l1, l1)
        bd = bd2 = 2 ** 31 - 1
        bi = -1
        for i2 in cand:
            d = dist(f1["desc1"][i1], f2.desc[i2])
            if md[i2] <= d:
                continue
            if d < bd:
                bd2, bd, bi = bd, d, i2
            elif d < bd2:
                bd2 = d
        if bd <= TH_LOW and F(bd) < F(F(bd2) * F(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1
                n -= 1
            m12[i1], m21[bi], md[bi] = bi, i1, bd
            n += 1
            if check_ori:
                hist[rot_bin(f1["angle1"][i1], f2.angle[bi])].append(i1)
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for i1 in hist[b]:
                    if m12[i1] >= 0:
                        m12[i1] = -1
                        n -= 1
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1] = (f2.x[m12[i1]], f2.y[m12[i1]])
    return n, np.array(m12, np.int32), prev
