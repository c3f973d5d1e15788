"""CPU oracle: numpy restatement of the gcmiipy Matsuno C-grid hot path, in SI magnitudes.

TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared with; it is
never imported by the product (`gcmiipy_b200/`).  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Every function restates one reference function (cited as file:line into
/root/reference) on plain float64 ndarrays, keeping the reference's operation order so that
the result is bit-identical to the reference executed through the SI pint stand-in
(`oracle/shims/pint`).  PINNING: `tests/test_oracle_pinning.py` checks this file against
(a) the reference's own known-answer tests (test_matsumo.py:9-29, flux_limiter.py:46-48,
temperature.py:31-41) and (b) `tests/golden/*.npz`, outputs of the UNMODIFIED reference
produced by `oracle/make_golden.py` in the build container.

Layout: 2-D fields [j, i]; 3-D fields [k, j, i]; i fastest; row 0 = north; k = 0 = surface.
"""
import math

import numpy as np

# ---- constants.py:16-48 (SI) -------------------------------------------------------------
Rd = 287.0
Cp = 1004.0
kappa = Rd / Cp            # constants.py:28
P0 = 100000.0              # constants.py:31
G = 9.8                    # constants.py:45
radius = 6.3781e6          # constants.py:48
mu_air = 18.5 * 1e-6      # constants.py:51  (18.5 uPa s)
Rv = 461.0                 # constants.py:78
standard_pressure = 101325.0
standard_temperature = 273.16


# ---- coordinates_3d.py:32-98 / coordinates.py:29-79 / coordinates_1d.py:25-53 -------------
def ipj(q): return np.roll(q, -1, -1)          # i+1
def imj(q): return np.roll(q, 1, -1)           # i-1
def ijp(q): return np.roll(q, -1, -2)          # j+1
def ijm(q): return np.roll(q, 1, -2)           # j-1
def imjp(q): return imj(ijp(q))
def kp(q): return np.roll(q, -1, -3)           # k+1
def km(q): return np.roll(q, 1, -3)            # k-1
def kph(q): return (q + kp(q)) / 2
def kmh(q): return (q + km(q)) / 2
def iph(q): return (q + ipj(q)) / 2
def imh(q): return (q + imj(q)) / 2
def jph(q): return (q + ijp(q)) / 2
def jmh(q): return (q + ijm(q)) / 2
def gradi(q, dx): return (ipj(q) - q) / dx
def gradj(q, dy): return (ijp(q) - q) / dy


# 1-D (coordinates_1d.py): arrays [i], roll along axis 0
def ip1(q): return np.roll(q, -1, 0)
def im1(q): return np.roll(q, 1, 0)


# ---- temperature.py:7-19 ---------------------------------------------------------------------
def to_true_temp(t, p):
    return t / ((P0 / p) ** kappa)


def to_potential_temp(tt, p):
    return tt * ((P0 / p) ** kappa)


# ---- geometry.py:9-182 -----------------------------------------------------------------------
class Geom:
    """geometry.py:9-26; all members are SI magnitudes with the reference's broadcast shapes."""
    pass


def manabe_sig(s):                      # geometry.py:30
    return s ** 2 * (3 - 2 * s)


def equal_sig(s):                       # geometry.py:34
    return s


def _sigma_tables(geom, layers, sig_func):   # geometry.py:73-85, :158-172
    mysig = [sig_func(1 - i / layers) for i in range(layers + 1)]
    rs = lambda a: np.reshape(np.asarray(a, dtype=float), (len(a), 1, 1))
    geom.sige = rs(mysig)
    geom.sigt = rs(mysig[1:])
    geom.sigb = rs(mysig[:-1])
    geom.dsig = geom.sigb - geom.sigt
    geom.sig = (geom.sigb + geom.sigt) / 2
    geom.dsigv = np.roll(geom.sig, -1, -3) - geom.sig


def gen_geometry(height, width, layers, sig_func=equal_sig,
                 north_edge=90, south_edge=-90, west_edge=-180, east_edge=180):
    """geometry.py:38-151 (prints dropped)."""
    geom = Geom()
    geom.height, geom.width, geom.layers = height, width, layers
    _sigma_tables(geom, layers, sig_func)
    circumference = 2 * radius * math.pi
    dlat = (north_edge - south_edge) / height
    dlong = (east_edge - west_edge) / width
    lat_j = np.zeros((height,))
    lat_h = np.zeros((height,))
    for i in range(height):
        lat_j[i] = north_edge - (i + 0.5) * dlat
        lat_h[i] = north_edge - (i + 1) * dlat
    long_k = np.zeros((width,))
    for i in range(width):
        long_k[i] = west_edge + (i + 0.5) * dlong
    geom.lat = lat_j.reshape((height, -1)) * (math.pi / 180.0)
    geom.long = long_k * (math.pi / 180.0)
    cos_j = np.cos(lat_j * np.pi / 180)
    cos_h = np.cos(lat_h * np.pi / 180)
    dx_j = cos_j * circumference / width
    dx_h = cos_h * circumference / width
    geom.dx_j = np.reshape(dx_j, (1, height, 1))
    geom.dx_h = np.reshape(dx_h, (1, height, 1))
    geom.dy = circumference / 2 / height
    geom.area = (np.roll(dx_h, 1, axis=0) + dx_h) * geom.dy * 0.5
    geom.ptop = 0.0
    geom.heightmap = np.zeros((height, width))
    return geom


def gen_square_geometry(height, width, layers, dx, dy, sig_func=equal_sig):
    """geometry.py:154-182."""
    geom = Geom()
    geom.height, geom.width, geom.layers = height, width, layers
    geom.ptop = 0.0
    _sigma_tables(geom, layers, sig_func)
    geom.lat = 0.0
    geom.long = 0.0
    geom.dx_j = np.full((1, height, 1), float(dx))
    geom.dx_h = np.full((1, height, 1), float(dx))
    geom.dy = float(dy)
    geom.heightmap = np.zeros((height, width))
    return geom


def pressure_from_heightmap(height, sea_level_pressure, sea_level_temp):
    """geometry.py:185-231 (returns `wiki_val`, :227): R = 8.3145 J/(K mol), Md = 28.97 g/mol (constants.py:10,13)."""
    R, Md = 8.3145, 28.97e-3
    return sea_level_pressure * np.exp((-G * Md * height) / (R * sea_level_temp))


# ---- low_pass.py:14-78 -----------------------------------------------------------------------
def polar_filter_table(geom, im):
    """smmz[..., n] of low_pass.py:61-72: 1 for n = 0, min(1, (dx_j/dy)/sin(pi n/im)) for n >= 1."""
    drat = geom.dy / geom.dx_j
    nmax = im / 2
    bysn = 1 / np.sin(np.pi / im * np.arange(1, nmax + 1))
    sm = 1 - bysn / drat
    smmz = 1 - np.maximum(sm, np.zeros_like(sm))
    return np.insert(smmz, 0, 1, -1)


def arakawa_1977(q, geom):
    """low_pass.py:41-78.  NB a 2-D input comes back as (1, H, W) because dx_j is (1, H, 1)."""
    im = q.shape[-1]
    if im == 1:
        return q
    smmz = polar_filter_table(geom, im)
    f_q = np.fft.rfft(q)
    return np.fft.irfft(f_q * smmz)


def avrx(q, geom):
    """low_pass.py:14-38: hard spectral cut-off, 2-D input only."""
    (jm, im) = q.shape
    f_q = np.fft.rfft(q)
    ratios = np.fft.rfftfreq(im, geom.dx_j) * geom.dy
    ratio_mult = np.zeros_like(ratios)
    ratio_mult[ratios <= 0.5] = 1
    return np.fft.irfft(f_q * ratio_mult)


# ---- dynamics.py:15-237 ----------------------------------------------------------------------
def calc_pu(p, u): return u * iph(p)            # dynamics.py:15
def calc_pv(p, v): return v * jph(p)            # dynamics.py:20
def un_pu(pu, p): return pu / iph(p)            # dynamics.py:25
def un_pv(pv, p): return pv / jph(p)            # dynamics.py:30


def aflux(pu, pv, geom):
    """dynamics.py:35-46."""
    conv = ((pu - imj(pu)) / geom.dx_j + (pv - ijm(pv)) / geom.dy) * geom.dsig
    pit = np.sum(conv, 0)
    sd = np.cumsum(conv[::-1], 0)[::-1] - pit * geom.sigb
    sd[0] = 0
    return pit, sd


def advec_sig(sd, q, geom):
    """dynamics.py:49-52."""
    flux = kmh(q) * sd
    dq = (flux - kp(flux)) / geom.dsig
    return -dq


def advec_m_pu(p, u, v, pu, pv, geom):
    """dynamics.py:55-108 (Coriolis hard-disabled at :82, adds exactly 0)."""
    puum = imh(u) * imh(pu)
    puup = ipj(puum)
    puvp = iph(pv) * jph(u)
    puvm = ijm(puvp)
    pvvm = jmh(v) * jmh(pv)
    pvvp = ijp(pvvm)
    pvup = iph(v) * jph(pu)
    pvum = imj(pvup)
    coriolis = 0.0
    dut = (puum - puup) / geom.dx_j + (puvm - puvp) / geom.dy + coriolis
    dvt = (pvvm - pvvp) / geom.dy + (pvum - pvup) / geom.dx_h + coriolis
    return dut, dvt


def compute_geopotential(p, t, geom):
    """dynamics.py:111-142 (the print-only `phi_mine` branch :116-119,:137-140 is dropped)."""
    tp = p * geom.sig + geom.ptop
    tt = to_true_temp(t, tp)
    rho = tp / (Rd * tt)
    sp = geom.sig * p
    spa = sp / rho
    s1 = spa * geom.dsig
    pkdn = ((geom.sig * p + geom.ptop) / P0) ** kappa
    pkup = kp(pkdn)
    stp = Cp * kph(t) * (pkdn - pkup)
    s2 = geom.sigt * stp
    stp_n = km(stp)
    stp_n[0] = np.sum(s1 - s2, 0) + geom.heightmap * G
    return np.cumsum(stp_n, 0)


def pgf(p, t, geom):
    """dynamics.py:147-171."""
    tp = p * geom.sig + geom.ptop
    tt = to_true_temp(t, tp)
    rho = tp / (Rd * tt)
    sp = geom.sig * p
    phi = compute_geopotential(p, t, geom)
    phiu = iph(p) * gradi(phi, geom.dx_j)
    phiv = jph(p) * gradj(phi, geom.dy)
    pgfu = iph(sp) / iph(rho) * gradi(p, geom.dx_j)
    pgfv = jph(sp) / jph(rho) * gradj(p, geom.dy)
    return pgfu, pgfv, phiu, phiv


def advec_t(pu, pv, t, geom):
    """dynamics.py:174-181."""
    tpu = pu * iph(t)
    tpv = pv * jph(t)
    return (tpu - imj(tpu)) / geom.dx_j + (tpv - ijm(tpv)) / geom.dy


def half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, geom):
    """dynamics.py:183-227."""
    pu = calc_pu(p, u)
    spu_orig = calc_pu(sp, su)
    spu = arakawa_1977(spu_orig, geom)
    pv = calc_pv(p, v)
    spv = calc_pv(sp, sv)
    pit, sd = aflux(spu, spv, geom)
    p_n = p - pit * dt
    dut, dvt = advec_m_pu(sp, su, sv, spu, spv, geom)
    pgu, pgv, phiu, phiv = pgf(sp, st, geom)
    dus = advec_sig(iph(sd), su, geom)
    dvs = advec_sig(jph(sd), sv, geom)
    pgfu = arakawa_1977(pgu + phiu, geom)
    pu_n = pu - (dut + dus + pgfu) * dt
    pv_n = pv - (dvt + dvs + phiv + pgv) * dt
    u_n = un_pu(pu_n, p_n)
    v_n = un_pv(pv_n, p_n)
    t_n = (t * p - (advec_t(spu, spv, st, geom) + advec_sig(sd, st, geom)) * dt) / p_n
    q_n = (q * p - (advec_t(spu, spv, sq, geom) + advec_sig(sd, sq, geom)) * dt) / p_n
    v_n[:, -1, :] *= 0
    return p_n, u_n, v_n, t_n, q_n


def matsuno_timestep(p, u, v, t, q, dt, geom, boundary_conditions=None):
    """dynamics.py:230-237."""
    sp, su, sv, st, sq = half_timestep(p, u, v, t, q, p, u, v, t, q, dt, geom)
    if boundary_conditions:
        sp, su, sv, st, sq = boundary_conditions(sp, su, sv, st, sq, dt, geom)
    op, ou, ov, ot, oq = half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, geom)
    if boundary_conditions:
        op, ou, ov, ot, oq = boundary_conditions(op, ou, ov, ot, oq, dt, geom)
    return op, ou, ov, ot, oq


# ---- humidity.py:4-31 + no_limits_2_5d.py:35-60,146-168,220-236 ---------------------------
def manabe_rh(geom):                                   # humidity.py:4
    return 0.77 * (geom.sig - 0.02) / (1 - 0.02)


def saturation_vapor_pressure(tt):                     # humidity.py:10 (Buck), Pa
    t = tt - 273.15
    return 0.61121 * 1e3 * np.exp((18.678 - t / 234.5) * (t / (257.14 + t)))


def rh_to_mmr(rh, tp, tt):                             # humidity.py:27
    e_s = saturation_vapor_pressure(tt)
    e = rh * e_s
    w = e * Rd / (Rv * (tp - e))
    return w / (w + 1)


def gen_initial_conditions(geom):
    """no_limits_2_5d.py:146-168, prognostic part (ground variables are outside the path)."""
    full = (geom.layers, geom.height, geom.width)
    surface = (geom.height, geom.width)
    p = np.full(surface, 1) * 100000.0 - geom.ptop
    u = np.full(full, 1) * 1.0
    v = np.full(full, 1) * .0
    tt = np.full(full, 1) * 360.0
    tp = p * geom.sig + geom.ptop
    t = to_potential_temp(tt, tp)
    q = np.full(full, 1) * 0.000003
    q = np.maximum(q, rh_to_mmr(manabe_rh(geom), tp, tt))
    return p, u, v, t, q


def run_model_ic(geom):
    """Initial state of no_limits_2_5d.run_model (:222-226): ICs, v[0,0,0]=0.1, u*=0."""
    p, u, v, t, q = gen_initial_conditions(geom)
    v[0, 0, 0] = 0.1
    u *= 0
    return p, u, v, t, q


def calc_energy(p, u, v, t, q, geom):
    """no_limits_2_5d.py:35-60.  geom.area is (H,) and broadcasts along i (reference quirk):
    only valid when H == W or H == 1."""
    mag = np.sqrt(imh(u) ** 2 + jmh(v) ** 2)
    tp = p * geom.sig + geom.ptop
    tt = to_true_temp(t, tp)
    rho = tp / (Rd * tt)
    dp = p * geom.dsig
    depth = dp / (rho * G)
    airmass = rho * depth * geom.area
    total_depth = np.cumsum(depth, 0)
    geo = np.sum(total_depth * airmass * G)
    ke = np.sum(mag ** 2 * .5 * airmass)
    ate = np.sum(tt * Cp * airmass)
    return ke, ate, geo, ke + ate + geo


# ---- matsuno_c_grid.py:15-142 ----------------------------------------------------------------
def advection_of_velocity_u(u, v, dx):
    """matsuno_c_grid.py:15-51."""
    u_ipj = (ipj(u) + u) / 2
    u_imj = (imj(u) + u) / 2
    v_ijm = (imj(v) + v) / 2
    v_ijp = (imjp(v) + ijp(v)) / 2
    du_ipj = (ipj(u) - u)
    du_imj = (u - imj(u))
    du_ijp = (ijp(u) - u)
    du_ijm = (u - ijm(u))
    return (u_ipj * du_ipj + u_imj * du_imj + v_ijp * du_ijp + v_ijm * du_ijm) / dx


def advection_of_velocity_v(u, v, dx):
    """matsuno_c_grid.py:54-80."""
    v_ijp = (ijp(v) + v) / 2
    v_ijm = (ijm(v) + v) / 2
    u_ipj = (u + ijm(u)) / 2
    u_imj = (imj(u) + imjp(u)) / 2
    dv_ipj = (ipj(v) - v)
    dv_imj = (v - imj(v))
    dv_ijp = (ijp(v) - v)
    dv_ijm = (v - ijm(v))
    return (u_ipj * dv_ipj + u_imj * dv_imj + v_ijp * dv_ijp + v_ijm * dv_ijm) / dx


def geopotential_gradient_u(p, dx):              # matsuno_c_grid.py:97
    return (ipj(p) - p) / dx * G


def geopotential_gradient_v(p, dx):              # matsuno_c_grid.py:103
    return (ijp(p) - p) / dx * G


def advection_of_geopotential(u, v, p, dx):
    """matsuno_c_grid.py:109-118."""
    up_imj = (imj(p) + p) / 2 * imj(u)
    up_ipj = (ipj(p) + p) / 2 * u
    vp_ijm = (ijm(p) + p) / 2 * ijm(v)
    vp_ijp = (ijp(p) + p) / 2 * v
    return (up_ipj - up_imj) / dx + (vp_ijp - vp_ijm) / dx


def courant_number(p, u, dx, dt):                # matsuno_c_grid.py:121
    return (np.max(u) + np.sqrt(np.mean(p) * G)) * dt / dx


def matsumo_scheme(u, v, p, dx, dt):
    """matsuno_c_grid.py:125-142."""
    u_s = u - dt * (advection_of_velocity_u(u, v, dx) + geopotential_gradient_u(p, dx))
    v_s = v - dt * (advection_of_velocity_v(u, v, dx) + geopotential_gradient_v(p, dx))
    p_s = p - dt * advection_of_geopotential(u, v, p, dx)
    u_n = u - dt * (advection_of_velocity_u(u_s, v_s, dx) + geopotential_gradient_u(p_s, dx))
    v_n = v - dt * (advection_of_velocity_v(u_s, v_s, dx) + geopotential_gradient_v(p_s, dx))
    p_n = p - dt * advection_of_geopotential(u_s, v_s, p_s, dx)
    return u_n, v_n, p_n


# ---- no_limits_2d.py:21-131 ------------------------------------------------------------------
def pe2d_advec_p(pu, pv, dx):                    # no_limits_2d.py:41
    return (pu - imj(pu)) / dx + (pv - ijm(pv)) / dx


def pe2d_advec_m(p, u, v, dx):
    """no_limits_2d.py:47-73."""
    vph = iph(v)
    p_mid = iph(jph(p))
    puum = imh(u) ** 2 * p
    puup = ipj(puum)
    puvm = jmh(u) * ijm(vph) * ijm(p_mid)
    puvp = ipj(puvm)
    dut = (puum - puup) / dx + (puvm - puvp) / dx
    pvvm = jmh(v) ** 2 * p
    pvvp = ijp(pvvm)
    pvum = imj(p_mid) * imh(v) * imj(jph(u))
    pvup = ipj(pvum)
    dvt = (pvvm - pvvp) / dx + (pvum - pvup) / dx
    return dut, dvt


def pe2d_pgf(p, t, dx):
    """no_limits_2d.py:76-89."""
    ppih = iph(p)
    ttu = to_true_temp(iph(t), ppih)
    rhou = ppih / (Rd * ttu)
    pgfu = ppih / rhou * gradi(p, dx)
    ppjh = jph(p)
    ttv = to_true_temp(jph(t), ppjh)
    rhov = ppjh / (Rd * ttv)
    pgfv = ppjh / rhov * gradj(p, dx)
    return pgfu, pgfv


def pe2d_advec_t(pu, pv, t, dx):                 # no_limits_2d.py:92
    tpu = pu * iph(t)
    tpv = pv * jph(t)
    return (tpu - imj(tpu)) / dx + (tpv - ijm(tpv)) / dx


def pe2d_half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, dx):
    """no_limits_2d.py:104-126 (q passes through unchanged)."""
    pu = calc_pu(p, u)
    spu = calc_pu(sp, su)
    pv = calc_pv(p, v)
    spv = calc_pv(sp, sv)
    p_n = p - pe2d_advec_p(spu, spv, dx) * dt
    dut, dvt = pe2d_advec_m(sp, su, sv, dx)
    pgu, pgv = pe2d_pgf(sp, st, dx)
    pu_n = pu - (dut + pgu) * dt
    pv_n = pv - (dvt + pgv) * dt
    u_n = un_pu(pu_n, p_n)
    v_n = un_pv(pv_n, p_n)
    t_n = t - (pe2d_advec_t(spu, spv, st, dx) / p_n) * dt
    return p_n, u_n, v_n, t_n, q


def pe2d_matsuno_timestep(p, u, v, t, q, dt, dx):
    """no_limits_2d.py:129-131."""
    sp, su, sv, st, sq = pe2d_half_timestep(p, u, v, t, q, p, u, v, t, q, dt, dx)
    return pe2d_half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, dx)


# ---- phi_port.py:5-136 -----------------------------------------------------------------------
def phi_port_PGF(T, P, geom):
    """phi_port.py:5-113.  T[W,H,L], P[W,H] (transposed views); only I = 0 is computed per J
    (IMAX = 1, :50-54); arithmetic-mean THETA (:78); EXPBYK(X) = X**kappa (:116)."""
    IM, JM, LM = T.shape
    SIG = geom.sig.flatten()
    DSIG = geom.dsig.flatten()
    SIGE = geom.sige.flatten()
    FDATA = np.transpose(geom.heightmap)
    SHA = Rd / kappa
    PHI = np.zeros_like(T)
    # the J loop is restated as whole-column array arithmetic over J (identical per-element
    # operations; array `**` is also what the reference's 0-d Quantities go through)
    I = 0
    SUM1 = np.zeros(JM)
    SUM2 = np.zeros(JM)
    SP = P[I, :]
    PDN = SIG[0] * SP + geom.ptop
    PKDN = PDN ** kappa
    for L in range(LM - 1):
        SPA = SIG[L] * SP * Rd * T[I, :, L] * PKDN / PDN
        SUM1 = SUM1 + SPA * DSIG[L]
        PUP = SIG[L + 1] * SP + geom.ptop
        PKUP = PUP ** kappa
        THETA = (T[I, :, L + 1] + T[I, :, L]) / 2
        PHI[I, :, L + 1] = SHA * THETA * (PKDN - PKUP)
        SUM2 = SUM2 + SIGE[L + 1] * PHI[I, :, L + 1]
        PDN = PUP
        PKDN = PKUP
    SPA = SIG[LM - 1] * SP * Rd * T[I, :, LM - 1] * PKDN / PDN
    SUM1 = SUM1 + SPA * DSIG[LM - 1]
    PHI[I, :, 0] = FDATA[I, :] + SUM1 - SUM2
    for L in range(1, LM):
        PHI[I, :, L] = PHI[I, :, L] + PHI[I, :, L - 1]
    return PHI


def THBAR(X, Y):                                  # phi_port.py:120-136
    x = X / Y
    return X * (np.log(x) / (x - 1))


# ---- viscosity.py:12-25 ----------------------------------------------------------------------
def finite_laplacian_2d(q, dx):
    top = ijp(q) + ijm(q) + ipj(q) + imj(q) - 4 * q
    return top / (dx * dx)


def incompressible_viscosity_2d(u, mu, dx):
    return mu * finite_laplacian_2d(u, dx)


# ---- flux_limiter.py:10-32 -------------------------------------------------------------------
def van_leer(r):
    return (r + np.abs(r)) / (1 + np.abs(r))


def calc_r(q):
    a = q - im1(q)
    b = ip1(q) - q
    return np.divide(a, b, out=np.zeros_like(a), where=(b != 0))


def donor_cell_flux(q, u):
    return np.where(u > 0, q, ip1(q)) * u


def donor_cell_advection(q, u, dx, dt):
    flux = donor_cell_flux(q, u)
    return q + (im1(flux) - flux) * dt / dx


# ---- matsumo_temp.py:13-99 (shallow water + temperature + viscosity) -------------
def mt_density_from(p, t):                        # matsumo_temp.py:13
    temp = t / ((100000.0 / p) ** (Rd / Cp))
    return p / (Rd * temp)


def mt_matsumo_scheme(u, v, p, t, dx, dt):
    """matsumo_temp.py:66-99 (keeps the u-for-v viscosity quirk at :75,:91)."""
    def tend(uu, vv, pp, tt):
        density = mt_density_from(pp, tt)
        geo = pp / (G * density)
        fu = (advection_of_velocity_u(uu, vv, dx) + geopotential_gradient_u(geo, dx)
              - incompressible_viscosity_2d(uu, mu_air, dx) / density)
        fv = (advection_of_velocity_v(uu, vv, dx) + geopotential_gradient_v(geo, dx)
              - incompressible_viscosity_2d(uu, mu_air, dx) / density)
        return fu, fv
    scaled_t = p * t * dx * dx
    fu, fv = tend(u, v, p, t)
    u_s = u - dt * fu
    v_s = v - dt * fv
    p_s = p - dt * advection_of_geopotential(u, v, p, dx)
    tt = scaled_t - dt * advection_of_geopotential(u, v, scaled_t, dx)
    t_s = tt / (p_s * dx * dx)
    scaled_t_s = p_s * t_s * dx * dx
    fu, fv = tend(u_s, v_s, p_s, t_s)
    u_n = u - dt * fu
    v_n = v - dt * fv
    p_n = p - dt * advection_of_geopotential(u_s, v_s, p_s, dx)
    tt_n = scaled_t - dt * advection_of_geopotential(u_s, v_s, scaled_t_s, dx)
    t_n = tt_n / (p_n * dx * dx)
    return u_n, v_n, p_n, t_n


# ---- opt-in terms of the 2.5-D half step ----------------------------------------
# The reference only sketches these (dead Coriolis branch dynamics.py:82-95; "TODO might need to flux limit
# this" dynamics.py:217-218; viscosity.py is wired into matsumo_temp.py only).  They are composed here from the
# reference's own building blocks.  PINNING: the Coriolis term is pinned against the reference executed with its
# `if False:` (dynamics.py:82) switched to `if True:` in memory (oracle/make_golden_ext.py ->
# tests/golden/run25_24x36x9_coriolis.npz).  The limiter and viscosity compositions have no reference
# counterpart: PARITY UNPINNED for those two; tests check them through their properties (conservation,
# monotonicity, reduction to the reference's operators).
class StepOptions:
    """coriolis: add dynamics.py:86-95; nu: kinematic horizontal viscosity (m2/s) on u and v; limit_q / limit_t:
    van Leer flux-limited horizontal advection of q / theta instead of the centred flux of advec_t."""

    def __init__(self, coriolis=False, nu=0.0, limit_q=False, limit_t=False):
        self.coriolis, self.nu, self.limit_q, self.limit_t = bool(coriolis), float(nu), bool(limit_q), bool(limit_t)

    def any(self):
        return self.coriolis or self.nu != 0.0 or self.limit_q or self.limit_t


def coriolis_parameters(geom):
    """(cp_at_u, cp_at_v) of dynamics.py:89-92, shapes (H, 1): 2 sin(lat) w at the u rows and at the v rows."""
    w = 2 * math.pi / 86400.0                                   # dynamics.py:89 (units.day)
    if np.ndim(geom.lat) == 0:                                  # gen_square_geometry: lat = 0 (geometry.py:174)
        z = np.zeros((geom.height, 1))
        return z, z
    cp_at_u = 2 * np.sin(geom.lat) * w                          # :91
    cp_at_v = 2 * np.sin(jph(geom.lat)) * w                     # :92
    return cp_at_u, cp_at_v


def coriolis_terms(pu, pv, geom):
    """dynamics.py:86-95 with the branch enabled -> (coriolis_u, coriolis_v)."""
    pu_at_pv = imh(jph(pu))                                     # :86
    pv_at_pu = iph(jmh(pv))                                     # :87
    cp_at_u, cp_at_v = coriolis_parameters(geom)
    return cp_at_u * -pv_at_pu, cp_at_v * pu_at_pv              # :94-95


def laplacian_h(q, dx, dy):
    """viscosity.py:12-19 on the lat-lon metric: second differences in i over dx^2 and in j over dy^2
    (identical to finite_laplacian_2d up to round-off when dx == dy)."""
    return (ipj(q) + imj(q) - 2 * q) / (dx * dx) + (ijp(q) + ijm(q) - 2 * q) / (dy * dy)


def limited_edge_value(q, flux, axis):
    """Edge value of q at i+1/2 (along `axis`) for a flux-limited scheme built from flux_limiter.py: the donor
    cell (:23-27, upwind by the sign of the flux) plus half the van Leer-limited (:10) downwind difference, the
    slope ratio being calc_r's (:14-20, 0 where the denominator is 0) seen from the upwind cell.
    phi = 1 gives the centred value iph(q) of advec_t, phi = 0 the donor cell."""
    sh = lambda a, n: np.roll(a, n, axis)
    b = sh(q, -1) - q                                                        # calc_r's b: q[i+1] - q[i]
    a = q - sh(q, 1)                                                         # calc_r's a: q[i] - q[i-1]
    r_up = np.divide(a, b, out=np.zeros_like(b), where=(b != 0))             # upwind cell i    (flux > 0)
    r_dn = np.divide(sh(b, -1), b, out=np.zeros_like(b), where=(b != 0))     # upwind cell i+1  (flux <= 0)
    return np.where(flux > 0, q + 0.5 * van_leer(r_up) * b, sh(q, -1) - 0.5 * van_leer(r_dn) * b)


def advec_t_limited(pu, pv, t, geom):
    """advec_t (dynamics.py:174-181) with iph(t), jph(t) replaced by the limited edge values."""
    tpu = pu * limited_edge_value(t, pu, -1)
    tpv = pv * limited_edge_value(t, pv, -2)
    return (tpu - imj(tpu)) / geom.dx_j + (tpv - ijm(tpv)) / geom.dy


def half_timestep_ext(p, u, v, t, q, sp, su, sv, st, sq, dt, geom, opt):
    """half_timestep (dynamics.py:183-227) with the opt-in terms; opt all off == half_timestep bit for bit."""
    pu = calc_pu(p, u)
    spu_orig = calc_pu(sp, su)
    spu = arakawa_1977(spu_orig, geom)
    pv = calc_pv(p, v)
    spv = calc_pv(sp, sv)
    pit, sd = aflux(spu, spv, geom)
    p_n = p - pit * dt
    dut, dvt = advec_m_pu(sp, su, sv, spu, spv, geom)
    if opt.coriolis:
        cu, cv = coriolis_terms(spu, spv, geom)
        dut = dut + cu                                                       # dynamics.py:100-101
        dvt = dvt + cv
    pgu, pgv, phiu, phiv = pgf(sp, st, geom)
    dus = advec_sig(iph(sd), su, geom)
    dvs = advec_sig(jph(sd), sv, geom)
    pgfu = arakawa_1977(pgu + phiu, geom)
    fu = dut + dus + pgfu
    fv = dvt + dvs + phiv + pgv
    if opt.nu != 0.0:     # pi * nu * lap(u): the mass-weighted form of matsumo_temp.py:55-59's mu * lap(u) / rho
        fu = fu - iph(sp) * (opt.nu * laplacian_h(su, geom.dx_j, geom.dy))
        fv = fv - jph(sp) * (opt.nu * laplacian_h(sv, geom.dx_h, geom.dy))
    pu_n = pu - fu * dt
    pv_n = pv - fv * dt
    u_n = un_pu(pu_n, p_n)
    v_n = un_pv(pv_n, p_n)
    adv_t = advec_t_limited if opt.limit_t else advec_t
    adv_q = advec_t_limited if opt.limit_q else advec_t
    t_n = (t * p - (adv_t(spu, spv, st, geom) + advec_sig(sd, st, geom)) * dt) / p_n
    q_n = (q * p - (adv_q(spu, spv, sq, geom) + advec_sig(sd, sq, geom)) * dt) / p_n
    v_n[:, -1, :] *= 0
    return p_n, u_n, v_n, t_n, q_n


def matsuno_timestep_ext(p, u, v, t, q, dt, geom, opt):
    """matsuno_timestep (dynamics.py:230-237) over half_timestep_ext."""
    s = half_timestep_ext(p, u, v, t, q, p, u, v, t, q, dt, geom, opt)
    return half_timestep_ext(p, u, v, t, q, *s, dt, geom, opt)


# ---- synthetic initial states shared by tests and bench ---------------
def synthetic_state(geom, seed=1234, amp_u=1.0, amp_p=50.0, amp_t=0.5):
    """Reference ICs (run_model_ic) plus a smooth seeded band-limited perturbation so that no
    term of the step is identically zero; v[:, -1, :] = 0 like every state the step produces."""
    p, u, v, t, q = run_model_ic(geom)
    L, H, W = u.shape
    rng = np.random.default_rng(seed)

    def smooth(shape):
        f = np.fft.rfft2(rng.standard_normal(shape))
        nj = max(1, shape[-2] // 8)
        ni = max(1, shape[-1] // 8)
        f[..., nj + 1:shape[-2] - nj, :] = 0
        f[..., :, ni + 1:] = 0
        x = np.fft.irfft2(f, s=shape[-2:])
        m = np.max(np.abs(x))
        return x / m if m > 0 else x

    u = u + amp_u * smooth((L, H, W))
    v = v + amp_u * smooth((L, H, W))
    p = p + amp_p * smooth((H, W))
    t = t + amp_t * smooth((L, H, W))
    v[:, -1, :] = 0
    return p, u, v, t, q


# ---- grey-radiation column physics: grey_solar.py:40-68, :323-333, :358-563; no_limits_2_5d.py:66-75 ----
sb_constant = 5.67e-8                 # W m-2 K-4    constants.py:71
solar_constant = 1.3608 * 1000.0      # W m-2        constants.py:59 (1.3608 kW m-2)
Cg = 1.13e6                           # J K-1 m-3    constants.py:25


def solar_zenith_angle(latitude, hour_angle, declination):
    """grey_solar.py:40-46: cosine of the solar zenith angle."""
    return np.sin(latitude) * np.sin(declination) + np.cos(latitude) * np.cos(declination) * np.cos(hour_angle)


def zenith_angle(longs, lats, utc_s, geom):
    """grey_solar.py:49-68 (declination 0): max(cos zenith, 0) [H, W]; utc_s = model time in seconds."""
    hour_angle = utc_s / (-24 * 3600.0) * 360 * (math.pi / 180.0)
    t_longs = np.tile(longs, (geom.height, 1))
    point_angle = t_longs + hour_angle
    return np.maximum(solar_zenith_angle(lats, point_angle, 0.0), 0)


def basic_grey_transmittances(t_lw, t_sw, geom):
    """grey_solar.py:323-333 (AD 2.35): per-layer long-wave / short-wave transmittance, shape of geom.dsig."""
    e_n = 1 - t_lw ** (geom.dsig)
    e_n_sw = 1 - t_sw ** (geom.dsig)
    return 1 - e_n, 1 - e_n_sw


def basic_grey_radiation(p, tp, tt, gt, t_lw, t_sw, albedo, utc_s, geom):
    """grey_solar.py:358-563 (the grey atmosphere of AD 2.7) -> (dT/dt [L, H, W] in K/s, d(ground T)/dt [H, W]).
    p: surface pressure, tp: layer pressure (unused by the reference too), tt: true temperature, gt: ground temperature."""
    L = geom.layers
    lw, sw = basic_grey_transmittances(t_lw, t_sw, geom)
    emission = (1 - lw) * sb_constant * tt ** 4
    cum_sw_from_top = np.cumprod(sw[::-1], axis=0)[::-1]
    cum_lw_from_bottom = np.cumprod(lw, axis=0)
    clw_b_div = cum_lw_from_bottom / lw
    B = np.sum(emission * clw_b_div, axis=0)                                   # 2.25
    sza = zenith_angle(geom.long, geom.lat, utc_s, geom)
    Sc = solar_constant * sza
    S = (1 - albedo) * Sc * cum_sw_from_top[0]                                 # 2.26
    U_s = 1 * sb_constant * gt ** 4                                            # 2.27
    dt_ground = (B + S - U_s) / Cg / 0.1
    shape = (L + 1, geom.height, geom.width)
    upwelling, downwelling = np.zeros(shape), np.zeros(shape)
    absorbed_dw = np.zeros(tt.shape)
    for i in reversed(range(L)):                                               # long wave from above
        absorbed_dw[i] = downwelling[i + 1] * (1 - lw[i])
        downwelling[i] = downwelling[i + 1] * lw[i] + emission[i]
    absorbed = np.zeros(tt.shape)
    for i in range(L):                                                         # long wave from below (atmosphere only)
        absorbed[i] = upwelling[i] * (1 - lw[i])
        upwelling[i + 1] = upwelling[i] * lw[i] + emission[i]
    U_n = clw_b_div * U_s * (1 - lw)                                           # 2.30
    S_n = (1 - sw) * cum_sw_from_top / sw * Sc                                 # 2.31
    B_n = emission                                                             # 2.32
    dTdt = (U_n + S_n - 2 * B_n + absorbed_dw + absorbed) * (G / (Cp * p * geom.dsig))   # 2.34
    return dTdt, dt_ground


def solar_timestep(t, p, gt, dt, utc_s, geom):
    """no_limits_2_5d.py:66-75 -> (theta_n, ground temperature_n); t_lw = 0.1, t_sw = 0.9, albedo = 0.3."""
    tp = p * geom.sig + geom.ptop
    tt = to_true_temp(t, tp)
    dt_air, dt_ground = basic_grey_radiation(p, tp, tt, gt, 0.1, 0.9, 0.3, utc_s, geom)
    gt_n = gt + dt_ground * dt
    tt_n = tt + dt_air * dt
    return to_potential_temp(tt_n, tp), gt_n
