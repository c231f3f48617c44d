//! C ABI of liblqft (include/lqft.h), transcribed 1:1.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct lqft_handle {
    _private: [u8; 0],
}

pub const LQFT_OK: i32 = 0;
pub const LQFT_ERR_INVALID: i32 = 1;
pub const LQFT_ERR_CUDA: i32 = 2;
pub const LQFT_ERR_NCCL: i32 = 3;
pub const LQFT_ERR_STATE: i32 = 4;
pub const LQFT_ERR_NOMEM: i32 = 5;
pub const LQFT_I32: u32 = 0;

#[repr(C)]
pub struct lqft_config {
    pub x: u32,
    pub y: u32,
    pub t: u32,
    pub dtype: u32,
    pub n_chains: u32,
    pub device: i32,
    pub world_size: u32,
    pub rank: u32,
    pub flags: u32,
    pub reserved: u32,
    pub nccl_id: [u8; 128],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct lqft_schedule {
    pub sweep_base: u64,
    pub first_step: u64,
    pub n_steps: u64,
    pub energy_every: u32,
    pub normalize_every: u32,
    pub action_every: u32,
    pub correlator_every: u32,
    pub correlator_reps: u32,
    pub difference_every: u32,
}

#[repr(C)]
pub struct lqft_run_outputs {
    pub e_obs: *mut f64,
    pub e_int: *mut i64,
    pub s_obs: *mut f64,
    pub s_int: *mut i64,
    pub corr: *mut f64,
    pub accepted: *mut u64,
}

extern "C" {
    pub fn lqft_version() -> u32;
    pub fn lqft_last_error() -> *const c_char;
    pub fn lqft_device_count() -> i32;
    pub fn lqft_comm_unique_id(id: *mut u8) -> i32;
    pub fn lqft_create(cfg: *const lqft_config, out: *mut *mut lqft_handle) -> i32;
    pub fn lqft_destroy(h: *mut lqft_handle) -> i32;
    pub fn lqft_set_chain(h: *mut lqft_handle, chain: u32, beta: f64, seed: u64, wilson_width: u32, wilson_height: u32) -> i32;
    pub fn lqft_init_hot(h: *mut lqft_handle) -> i32;
    pub fn lqft_upload(h: *mut lqft_handle, chain: u32, values: *const c_void) -> i32;
    pub fn lqft_download(h: *mut lqft_handle, chain: u32, values: *mut c_void) -> i32;
    pub fn lqft_sweep(h: *mut lqft_handle, first_sweep: u64, n_sweeps: u32, accepted: *mut u64) -> i32;
    pub fn lqft_normalize(h: *mut lqft_handle, sweep_index: u64) -> i32;
    pub fn lqft_normalize_site(h: *mut lqft_handle, chain: u32, site: u64) -> i32;
    pub fn lqft_shift_values(h: *mut lqft_handle, chain: u32, shift: f64) -> i32;
    pub fn lqft_energy(h: *mut lqft_handle, e_int: *mut i64, e_obs: *mut f64) -> i32;
    pub fn lqft_action(h: *mut lqft_handle, s_int: *mut i64, s_obs: *mut f64) -> i32;
    pub fn lqft_slice_moments(h: *mut lqft_handle, chain: u32, s1: *mut i64, s2: *mut i64) -> i32;
    pub fn lqft_correlator(h: *mut lqft_handle, chain: u32, ref_sites: *const u64, reps: u32, sweep_index: u64, c_of_tau: *mut f64) -> i32;
    pub fn lqft_difference_update(h: *mut lqft_handle, chain: u32) -> i32;
    pub fn lqft_difference_read(h: *mut lqft_handle, chain: u32, mean: *mut f64, count: *mut u64) -> i32;
    pub fn lqft_run_measurements(first_step: u64, n_steps: u64, every: u32) -> u64;
    pub fn lqft_run(h: *mut lqft_handle, s: *const lqft_schedule, e_obs: *mut f64, e_int: *mut i64, accepted: *mut u64) -> i32;
    pub fn lqft_run_observables(h: *mut lqft_handle, s: *const lqft_schedule, out: *const lqft_run_outputs) -> i32;
    pub fn lqft_synchronize(h: *mut lqft_handle) -> i32;
    pub fn lqft_last_device_ms(h: *mut lqft_handle, ms: *mut f32) -> i32;
    pub fn lqft_kernel_launches(h: *mut lqft_handle, n: *mut u64) -> i32;
    pub fn lqft_narrow_segments(h: *mut lqft_handle, n: *mut u64) -> i32;
    pub fn lqft_plan_info(h: *mut lqft_handle, buf: *mut c_char, cap: usize) -> i32;
    pub fn lqft_stream(h: *mut lqft_handle, stream: *mut *mut c_void) -> i32;
    pub fn lqft_host_philox4x32(seed: u64, ctr: *const u32, out: *mut u32);
    pub fn lqft_host_site_word(seed: u64, idx: u64, colour: u32, sweep: u64) -> u32;
    pub fn lqft_host_threshold_table(beta: f64, thr: *mut u32, capacity: u32) -> u32;
    pub fn lqft_host_pick_site(seed: u64, stream: u32, counter: u64, rep: u32, n_sites: u64) -> u64;
    pub fn lqft_host_hot_value(seed: u64, idx: u64) -> i32;
    pub fn lqft_host_index(x_ext: u32, y_ext: u32, t_ext: u32, x: u32, y: u32, t: u32) -> u64;
    pub fn lqft_host_coords(x_ext: u32, y_ext: u32, t_ext: u32, idx: u64, xyz: *mut u32);
    pub fn lqft_host_neighbours(x_ext: u32, y_ext: u32, t_ext: u32, idx: u64, nbr: *mut u64);
    pub fn lqft_host_wilson_offset(y_ext: u32, x: u32, y: u32, t: u32, width: u32, height: u32) -> i32;
    pub fn lqft_host_slab(t_ext: u32, world_size: u32, rank: u32, t0: *mut u32, nt: *mut u32) -> i32;
}
