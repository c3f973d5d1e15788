"""ctypes declarations of include/gcm_b200.h (the C ABI of libgcm_b200.so).  Declarations only."""
import ctypes as C

c_dp = C.c_void_p          # device pointer to float64
c_stream = C.c_void_p      # cudaStream_t


class GeomDesc(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("L", C.c_int), ("wrap_j", C.c_int), ("row_lo", C.c_int),
                ("row_hi", C.c_int), ("zero_v_row", C.c_int), ("dy", C.c_double), ("ptop", C.c_double),
                ("h_sig", C.c_void_p), ("h_dsig", C.c_void_p), ("h_sigb", C.c_void_p), ("h_sigt", C.c_void_p),
                ("h_dx_j", C.c_void_p), ("h_dx_h", C.c_void_p), ("h_heightmap", C.c_void_p), ("h_smmz", C.c_void_p),
                ("zero_v_row2", C.c_int)]


class Pe25Options(C.Structure):
    _fields_ = [("coriolis", C.c_int), ("limit_q", C.c_int), ("limit_t", C.c_int), ("nu", C.c_double),
                ("h_cor_u", C.c_void_p), ("h_cor_v", C.c_void_p)]


class State(C.Structure):
    _fields_ = [("p", c_dp), ("u", c_dp), ("v", c_dp), ("t", c_dp), ("q", c_dp)]


_geom = C.c_void_p
_st = C.POINTER(State)
_i, _d, _z = C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes)
SIGNATURES = {
    "gcm_version": (_i, []),
    "gcm_status_string": (C.c_char_p, [_i]),
    "gcm_last_status": (_i, [_i]),
    "gcm_nonfinite_read": (_i, [C.c_void_p, _i, c_stream]),
    "gcm_geom_create": (_i, [C.POINTER(GeomDesc), C.POINTER(_geom)]),
    "gcm_geom_destroy": (_i, [_geom]),
    "gcm_pe25_workspace_bytes": (_z, [_geom, _i]),
    "gcm_pe25_workspace_field": (_z, [_geom, _i, _i]),
    "gcm_pe25_half_step": (_i, [_geom, _st, _st, _st, _d, _i, c_dp, _z, c_stream]),
    "gcm_pe25_matsuno_step": (_i, [_geom, _st, _st, _d, _i, _i, c_dp, _z, c_stream]),
    "gcm_pe25_half_step_rows": (_i, [_geom, _st, _st, _st, _d, _i, c_dp, _z, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     c_stream]),
    "gcm_pe25_matsuno_step_host": (_i, [_geom, _st, _st, _st, _st, _st, _d, _i, c_dp, _z, c_stream]),
    "gcm_pe25_matsuno_step_host_pipelined": (_i, [_geom, _st, _st, _st, _st, _st, _d, _i, _i, c_dp, _z, c_stream]),
    "gcm_host_pipe_join": (_i, [c_stream]),
    "gcm_comm_unique_id": (_i, [C.c_void_p]),
    "gcm_comm_create": (_i, [_i, _i, C.c_void_p, C.POINTER(C.c_void_p)]),
    "gcm_comm_destroy": (_i, [C.c_void_p]),
    "gcm_comm_peer_setup": (_i, [C.c_void_p, _geom, C.c_void_p]),
    "gcm_comm_peer_connect": (_i, [C.c_void_p, C.c_void_p, C.c_void_p, _i]),
    "gcm_comm_peer_status": (_i, [C.c_void_p, C.POINTER(C.c_uint)]),
    "gcm_band_halo_peer": (_i, [_geom, C.c_void_p, _st, _i, _i, _i, c_stream]),
    "gcm_halo_exchange_begin": (_i, [_geom, C.c_void_p, _st, _i, _i, c_stream]),
    "gcm_halo_exchange_end": (_i, [_geom, C.c_void_p, _st, _i, _i, c_stream]),
    "gcm_band_matsuno_step": (_i, [_geom, C.c_void_p, _st, _st, _st, _d, _i, _i, c_dp, _z, c_stream]),
    "gcm_pe25_set_options": (_i, [_geom, C.POINTER(Pe25Options)]),
    "gcm_pe25_select_path": (_i, [_i]),
    "gcm_tuning_knob": (_i, [_i, _i]),
    "gcm_pe25_calc_pu": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_calc_pv": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_un_pu": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_un_pv": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_aflux": (_i, [_geom, c_dp, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_advec_sig": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_advec_m_pu": (_i, [_geom, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_geopotential": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_pe25_pgf": (_i, [_geom, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, _z, c_stream]),
    "gcm_pe25_advec_t": (_i, [_geom, c_dp, c_dp, c_dp, c_dp, c_stream]),
    "gcm_polar_filter": (_i, [_geom, c_dp, c_dp, _i, c_dp, c_stream]),
    "gcm_diag_minmax": (_i, [c_dp, _z, c_dp, c_stream]),
    "gcm_pe25_energy": (_i, [_geom, _st, c_dp, c_dp, c_stream]),
    "gcm_halo_buffer_doubles": (_z, [_geom, _i]),
    "gcm_halo_pack": (_i, [_geom, _st, _i, _i, c_dp, c_stream]),
    "gcm_halo_unpack": (_i, [_geom, _st, _i, _i, c_dp, c_stream]),
    "gcm_halo_copy_rows": (_i, [_geom, _st, _i, _st, _i, _i, c_stream]),
    "gcm_sw2d_matsuno_step": (_i, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, _i, _i, _d, _d, _i, c_dp, _z, c_stream]),
    "gcm_sw2d_workspace_bytes": (_z, [_i, _i]),
    "gcm_sw2d_operator": (_i, [_i, c_dp, c_dp, c_dp, c_dp, _i, _i, _d, c_stream]),
    "gcm_pe2d_half_step": (_i, [_st, _st, _st, _i, _i, _d, _d, c_stream]),
    "gcm_pe2d_matsuno_step": (_i, [_st, _st, _i, _i, _d, _d, _i, c_dp, _z, c_stream]),
    "gcm_pe2d_workspace_bytes": (_z, [_i, _i]),
    "gcm_pe2d_operator": (_i, [_i, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, _i, _i, _d, c_stream]),
    "gcm_swt2d_matsuno_step": (_i, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, _i, _i, _d, _d, _d, _i, c_dp, _z,
                                    c_stream]),
    "gcm_swt2d_workspace_bytes": (_z, [_i, _i]),
    "gcm_laplacian5": (_i, [c_dp, c_dp, _i, _i, _d, _d, _i, c_stream]),
    "gcm_phi_port_pgf": (_i, [_geom, c_dp, c_dp, c_dp, c_stream]),
    "gcm_fl_van_leer": (_i, [c_dp, c_dp, _z, c_stream]),
    "gcm_fl_calc_r": (_i, [c_dp, c_dp, _i, _i, c_stream]),
    "gcm_fl_donor_cell_flux": (_i, [c_dp, c_dp, c_dp, _i, _i, c_stream]),
    "gcm_fl_donor_cell_advection": (_i, [c_dp, c_dp, c_dp, _i, _i, _d, _d, _i, c_dp, c_stream]),
    "gcm_shift_op": (_i, [_i, c_dp, c_dp, _i, _i, _i, _i, _i, _d, c_stream]),
    "gcm_prof_enable": (_i, [_i]),
    "gcm_prof_kinds": (_i, []),
    "gcm_prof_kind_name": (C.c_char_p, [_i]),
    "gcm_prof_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "gcm_grey_radiation": (_i, [_geom, c_dp, c_dp, c_dp, C.c_void_p, C.c_void_p, _d, c_dp, c_dp, c_dp, _d, c_dp, c_dp,
                                c_stream]),
    "gcm_solar_timestep": (_i, [_geom, c_dp, c_dp, c_dp, C.c_void_p, C.c_void_p, _d, c_dp, c_dp, c_dp, _d, _d, c_dp, c_dp,
                                c_stream]),
    "gcm_temperature_convert": (_i, [_i, c_dp, c_dp, c_dp, _z, c_stream]),
}


def bind(cdll):
    """Attach restype/argtypes to every exported entry point; raises AttributeError if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll
