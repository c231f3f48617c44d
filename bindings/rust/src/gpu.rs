//! Owning wrapper over the liblqft handle plus the re-pointed call sites of the reference crate.
//! (`use crate::...` paths are the reference's own modules.)
use std::error::Error;
use std::ffi::CStr;

use crate::ffi::*;

/// `Box<dyn Error>` carrying lqft_last_error(), which the drivers already swallow with `.ok()` (sim.rs:92).
pub fn check(status: i32) -> Result<(), Box<dyn Error>> {
    if status == LQFT_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(lqft_last_error()) }.to_string_lossy().into_owned();
    Err(format!("lqft status {}: {}", status, msg).into())
}

/// The device-resident chains of one run; lives where `Simulation::run` owns its `Field` today.
pub struct Gpu {
    raw: *mut lqft_handle,
    pub n_chains: u32,
    pub t: usize,
    pub sites: usize,
}

impl Gpu {
    pub fn create(size: [usize; 3], n_chains: u32) -> Result<Self, Box<dyn Error>> {
        let cfg = lqft_config {
            x: size[0] as u32, y: size[1] as u32, t: size[2] as u32,
            dtype: LQFT_I32, n_chains, device: 0, world_size: 1, rank: 0, flags: 0, reserved: 0,
            nccl_id: [0u8; 128],
        };
        let mut raw = std::ptr::null_mut();
        check(unsafe { lqft_create(&cfg, &mut raw) })?;
        Ok(Gpu { raw, n_chains, t: size[2], sites: size[0] * size[1] * size[2] })
    }
    pub fn raw(&self) -> *mut lqft_handle { self.raw }
    pub fn set_chain(&self, chain: u32, temp: f64, seed: u64, width: usize, height: usize) -> Result<(), Box<dyn Error>> {
        check(unsafe { lqft_set_chain(self.raw, chain, temp, seed, width as u32, height as u32) })
    }
    pub fn init_hot(&self) -> Result<(), Box<dyn Error>> { check(unsafe { lqft_init_hot(self.raw) }) }
    /// Loop body of Simulation::run for `n_steps` steps; returns the energies (chain-major).
    pub fn run(&self, sweep_base: u64, first_step: u64, n_steps: u64, energy_every: u32, normalize_every: u32)
        -> Result<Vec<f64>, Box<dyn Error>> {
        let s = lqft_schedule { sweep_base, first_step, n_steps, energy_every, normalize_every, action_every: 0,
                                correlator_every: 0, correlator_reps: 0, difference_every: 0 };
        let n = unsafe { lqft_run_measurements(first_step, n_steps, energy_every) } as usize;
        let mut e = vec![0f64; n * self.n_chains as usize];
        let p = if n > 0 { e.as_mut_ptr() } else { std::ptr::null_mut() };
        check(unsafe { lqft_run(self.raw, &s, p, std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(e)
    }
    /// The whole measurement loop of `Simulation::run` (computation.rs:247-270) in one call: every `OutputData`
    /// kind on its own frequency, logged on the device.  Returns (energies, actions, correlator rows) of chain 0.
    pub fn run_observables(&self, s: &lqft_schedule) -> Result<(Vec<f64>, Vec<f64>, Vec<Vec<f64>>), Box<dyn Error>> {
        let nc = self.n_chains as usize;
        let ne = unsafe { lqft_run_measurements(s.first_step, s.n_steps, s.energy_every) } as usize;
        let na = unsafe { lqft_run_measurements(s.first_step, s.n_steps, s.action_every) } as usize;
        let ncr = unsafe { lqft_run_measurements(s.first_step, s.n_steps, s.correlator_every) } as usize;
        let (mut e, mut a, mut c) = (vec![0f64; ne * nc], vec![0f64; na * nc], vec![0f64; ncr * nc * self.t]);
        let nullable = |v: &mut Vec<f64>| if v.is_empty() { std::ptr::null_mut() } else { v.as_mut_ptr() };
        let out = lqft_run_outputs {
            e_obs: nullable(&mut e), e_int: std::ptr::null_mut(), s_obs: nullable(&mut a), s_int: std::ptr::null_mut(),
            corr: nullable(&mut c), accepted: std::ptr::null_mut(),
        };
        check(unsafe { lqft_run_observables(self.raw, s, &out) })?;
        e.truncate(ne);
        a.truncate(na);
        let rows = c[..ncr * self.t].chunks(self.t).map(|r| r.to_vec()).collect();
        Ok((e, a, rows))
    }
    pub fn sweep(&self, sweep: u64) -> Result<usize, Box<dyn Error>> {
        let mut acc = vec![0u64; self.n_chains as usize];
        check(unsafe { lqft_sweep(self.raw, sweep, 1, acc.as_mut_ptr()) })?;
        Ok(acc[0] as usize)
    }
    pub fn energy(&self) -> Result<f64, Box<dyn Error>> {
        let mut e = vec![0f64; self.n_chains as usize];
        check(unsafe { lqft_energy(self.raw, std::ptr::null_mut(), e.as_mut_ptr()) })?;
        Ok(e[0])
    }
    pub fn correlator(&self, reps: u64, sweep: u64) -> Result<Vec<f64>, Box<dyn Error>> {
        let mut c = vec![0f64; self.t];
        check(unsafe { lqft_correlator(self.raw, 0, std::ptr::null(), reps as u32, sweep, c.as_mut_ptr()) })?;
        Ok(c)
    }
    /// DifferencePlotOutputData::plot (outputdata.rs:258-276): mean[3][sites] of chain `chain` and the number of
    /// configurations accumulated by `difference_every` / lqft_difference_update.
    pub fn difference_read(&self, chain: u32) -> Result<(Vec<f64>, u64), Box<dyn Error>> {
        let mut mean = vec![0f64; 3 * self.sites];
        let mut count = 0u64;
        check(unsafe { lqft_difference_read(self.raw, chain, mean.as_mut_ptr(), &mut count) })?;
        Ok((mean, count))
    }
    pub fn download(&self, values: &mut [i32]) -> Result<(), Box<dyn Error>> {
        check(unsafe { lqft_download(self.raw, 0, values.as_mut_ptr() as *mut _) })
    }
}

impl Drop for Gpu {
    fn drop(&mut self) { unsafe { lqft_destroy(self.raw) }; }
}

// ---- call site 1: src/computation.rs, Simulation::run (replaces lines 221-281) -----------------------------
// fn run(mut self) -> Result<Computation<'a, D, SIZE>, Box<dyn Error>> {
//     let time = Instant::now();
//     let gpu = Gpu::create([self.lattice.size[0], self.lattice.size[1], self.lattice.size[2]], 1)?;
//     gpu.set_chain(0, self.temp, self.seed, 0, 0)?;
//     gpu.init_hot()?;                                             // Field::<i8>::random -> Field::<i32>::from_field
//     gpu.run(0, 0, self.burnin, 0, 10)?;                          // burn-in: sweeps + normalize every 10th step
//     let freq = self.output.iter().find_map(|o| o.energy_frequency()).unwrap_or(0) as u32;
//     let energies = gpu.run(self.burnin, 0, self.iterations, freq, 10)?;
//     self.output.iter_mut().for_each(|o| o.extend_energy(&energies));
//     self.duration = Some(time.elapsed().as_secs());
//     Ok(self.into_computation())
// }
//
// ---- call site 2: src/algorithm.rs, Metropolis::field_sweep (replaces 115-123) ---------------------------------
// fn field_sweep(&self, field: &mut Field<T, D, SIZE>, temp: f64, _rng: &mut ThreadRng) -> usize {
//     field.gpu.set_chain(0, temp, field.seed, 0, 0).unwrap();
//     field.gpu.sweep(field.next_sweep()).unwrap()
// }
//
// ---- call site 3: src/outputdata.rs, EnergyObservable::update (replaces 172-177) --------------------------------
// fn update<T, A: Action<T, D>>(&mut self, field: &A, _rng: &mut ThreadRng) {
//     self.values.push(field.gpu().energy().unwrap());              // == temp * float_energy(), fields.rs:669-671
// }
