//! B200 (sm_100a) backend for the batched PSF / FIPS 203 hot path of qfall-tools, behind the reference's own traits.
//!
//! * [`PSFGPVB200`] implements `qfall_tools::primitive::psf::PSF` (src/primitive/psf.rs:39-81) with the associated types
//!   of `PSFGPV` (gpv.rs:59-63), one target per call exactly like the reference, plus [`PSFBatch`] -- the extension the
//!   reference lacks because its methods reject multi-column inputs (gpv.rs:221, pinned by gpv.rs:287-298).
//! * [`PSFPerturbationB200`] (perturbation.rs) and [`PSFGPVRingB200`] (ring.rs) do the same for `PSFPerturbation`
//!   (mp_perturbation.rs:193-197) and `PSFGPVRing` (gpv_ring.rs:69-73).
//! * [`PolynomialRingZqB200`] / [`MatPolynomialRingZqB200`] (compression.rs) implement `LossyCompressionFIPS203`
//!   (lossy_compression_fips203.rs:20-59) over [`compress_words`] / [`decompress_words`], which run `Compress_d` /
//!   `Decompress_d` (:101-111, :159-169) for a whole coefficient vector at once.
//!
//! Values cross the boundary as fixed-width words: Range / key entries are residues in [0, q), q < 2^62 (`i64`), Domain
//! entries are `i32`.  Every C entry point returns a status; a non-zero status is turned into the panic (or the
//! `MathError`) the reference raises at the same place.
//!
//! This crate is shipped as SOURCE: the image the backend is built in has no Rust toolchain, so it has not been
//! compiled there.  The `extern "C"` block (ffi.rs) is checked mechanically against include/qfall_b200.h.
pub mod compression;
pub mod ffi;
pub mod perturbation;
pub mod ring;
pub use compression::{MatPolynomialRingZqB200, PolynomialRingZqB200};
pub use perturbation::PSFPerturbationB200;
pub use ring::PSFGPVRingB200;

use ffi::*;
use qfall_math::{
    integer::{MatZ, Z},
    integer_mod_q::MatZq,
    rational::{MatQ, Q},
    traits::{MatrixDimensions, MatrixGetEntry, MatrixSetEntry},
};
use qfall_tools::primitive::psf::{PSF, PSFGPV};
use std::cell::RefCell;
use std::ffi::CStr;

/// Owner of one `qf_ctx` (device memory, stream, installed key).  Neither `Send` nor `Sync`, like the reference's PSF
/// structs (gadget_parameters.rs:51): one caller thread per instance.
pub struct Context {
    pub(crate) raw: *mut qf_ctx,
}

impl Context {
    pub fn new(params: &qf_params, device: i32) -> Result<Self, String> {
        let mut raw: *mut qf_ctx = std::ptr::null_mut();
        let st = unsafe { qf_ctx_create(params, device, &mut raw) };
        if st != QF_OK || raw.is_null() {
            return Err(format!("qf_ctx_create failed with status {st}"));
        }
        Ok(Context { raw })
    }
    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(qf_last_error(self.raw)) }.to_string_lossy().into_owned()
    }
    /// status -> the reference's failure mode at that call site (`unwrap()` / `assert!`)
    pub(crate) fn check(&self, st: i32, what: &str) {
        assert!(st == QF_OK, "{what}: status {st}: {}", self.last_error());
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { qf_ctx_destroy(self.raw) }
    }
}

// ---- value conversion ------------------------------------------------------------------------------------------

/// Residues of a `MatZq` (least non-negative representatives), row-major.
pub fn matzq_to_words(m: &MatZq) -> Vec<i64> {
    let (rows, cols) = (m.get_num_rows(), m.get_num_columns());
    let mut out = Vec::with_capacity((rows * cols) as usize);
    for i in 0..rows {
        for j in 0..cols {
            let z: Z = m.get_entry(i, j).unwrap();
            out.push(i64::try_from(&z).expect("modulus below 2^62"));
        }
    }
    out
}

/// Entries of a `MatZ`, row-major.
pub fn matz_to_words(m: &MatZ) -> Vec<i64> {
    let (rows, cols) = (m.get_num_rows(), m.get_num_columns());
    let mut out = Vec::with_capacity((rows * cols) as usize);
    for i in 0..rows {
        for j in 0..cols {
            let z: Z = m.get_entry(i, j).unwrap();
            out.push(i64::try_from(&z).expect("entry below 2^63"));
        }
    }
    out
}

/// A Domain value (column vector) as `i32`; `None` when an entry does not fit -- such a vector is far outside D_n.
pub fn domain_to_i32(sigma: &MatZ) -> Option<Vec<i32>> {
    matz_to_words(sigma).into_iter().map(|v| i32::try_from(v).ok()).collect()
}

pub fn column_from_i32(e: &[i32]) -> MatZ {
    let mut out = MatZ::new(e.len() as i64, 1);
    for (i, v) in e.iter().enumerate() {
        out.set_entry(i as i64, 0, Z::from(*v as i64)).unwrap();
    }
    out
}

pub fn range_from_words(u: &[i64], q: &Z) -> MatZq {
    let mut out = MatZq::new(u.len() as i64, 1, q);
    for (i, v) in u.iter().enumerate() {
        out.set_entry(i as i64, 0, Z::from(*v)).unwrap();
    }
    out
}

fn matq_to_f64(m: &MatQ) -> Vec<f64> {
    let (rows, cols) = (m.get_num_rows(), m.get_num_columns());
    let mut out = Vec::with_capacity((rows * cols) as usize);
    for i in 0..rows {
        for j in 0..cols {
            let x: Q = m.get_entry(i, j).unwrap();
            out.push(f64::from(&x));
        }
    }
    out
}

// ---- the batch extension -----------------------------------------------------------------------------------------

/// Batched `samp_p` / `f_a` next to the unchanged trait: targets are independent given the key, samplers are keyed by
/// `(seed, first_index + position)`, so a batch split across calls or GPUs gives the same preimages as one call.
pub trait PSFBatch: PSF {
    fn samp_p_batch(&self, a: &Self::A, td: &Self::Trapdoor, us: &[Self::Range], seed: u64, first_index: u64) -> Vec<Self::Domain>;
    fn f_a_batch(&self, a: &Self::A, sigmas: &[Self::Domain]) -> Vec<Self::Range>;
}

// ---- PSFGPV ------------------------------------------------------------------------------------------------------

/// `PSFGPV` (gpv.rs:53-57) evaluated on the GPU.  `inner` keeps the reference struct (parameters, serde), `ctx` the
/// device context; the key and trapdoor most recently used are cached on the device.
pub struct PSFGPVB200 {
    pub inner: PSFGPV,
    ctx: Context,
    installed: RefCell<Option<(Vec<i64>, Vec<i64>)>>, // (A, S) words of the installed key
    seed: RefCell<u64>,
}

impl PSFGPVB200 {
    pub fn new(inner: PSFGPV, device: i32, seed: u64) -> Result<Self, String> {
        let gp = &inner.gp;
        let params = qf_params {
            kind: QF_PSF_GPV,
            n: i64::try_from(&gp.n).unwrap(),
            k: i64::try_from(&gp.k).unwrap(),
            m_bar: i64::try_from(&gp.m_bar).unwrap(),
            base: i64::try_from(&gp.base).unwrap(),
            q: u64::try_from(&Z::from(&gp.q)).map_err(|_| "modulus must be below 2^62".to_string())?,
            s: f64::from(&inner.s),
            r: 1.0,
            // gpv.rs:223: ||sigma||^2 <= s^2 * m, compared exactly (floor of the exact rational; 0 would make the
            // backend derive it from the f64-rounded s)
            norm_bound: ring::floor_u64(&(&inner.s * &inner.s * Q::from(i64::try_from(&(&gp.m_bar + &gp.n * &gp.k)).unwrap()))),
        };
        Ok(PSFGPVB200 { inner, ctx: Context::new(&params, device)?, installed: RefCell::new(None), seed: RefCell::new(seed) })
    }
    fn n(&self) -> usize {
        i64::try_from(&self.inner.gp.n).unwrap() as usize
    }
    fn m(&self) -> usize {
        (i64::try_from(&self.inner.gp.m_bar).unwrap() + i64::try_from(&self.inner.gp.n).unwrap() * i64::try_from(&self.inner.gp.k).unwrap()) as usize
    }
    /// a fresh sampler seed per call (the reference draws from its thread RNG; no seed API, SURVEY finding 1)
    fn next_seed(&self) -> u64 {
        let mut s = self.seed.borrow_mut();
        *s = s.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
        *s
    }
    /// qf_set_a + qf_set_trapdoor_gpv, skipped when the same key is already on the device
    fn install(&self, a: &MatZq, td: &(MatZ, MatQ)) {
        let (aw, sw) = (matzq_to_words(a), matz_to_words(&td.0));
        if let Some((a0, s0)) = self.installed.borrow().as_ref() {
            if *a0 == aw && *s0 == sw {
                return;
            }
        }
        self.ctx.check(unsafe { qf_set_a(self.ctx.raw, aw.as_ptr()) }, "qf_set_a");
        let gso = matq_to_f64(&td.1);
        self.ctx.check(unsafe { qf_set_trapdoor_gpv(self.ctx.raw, sw.as_ptr(), gso.as_ptr()) }, "qf_set_trapdoor_gpv");
        *self.installed.borrow_mut() = Some((aw, sw));
    }
}

impl PSF for PSFGPVB200 {
    type A = MatZq;
    type Trapdoor = (MatZ, MatQ);
    type Domain = MatZ;
    type Range = MatZq;

    /// gpv.rs:83-94 on the device: uniform A_bar and R, `A = [A_bar | G - A_bar R]` (qf_trap_gen), the short basis
    /// (qf_gen_short_basis, short_basis_classical.rs:54-110) and its GSO (qf_gso; fp64 where the reference is exact).
    fn trap_gen(&self) -> (MatZq, (MatZ, MatQ)) {
        let (n, m) = (self.n(), self.m());
        let m_bar = i64::try_from(&self.inner.gp.m_bar).unwrap() as usize;
        let (mut a, mut r) = (vec![0i64; n * m], vec![0i8; m_bar * (m - m_bar)]);
        self.ctx.check(unsafe { qf_trap_gen(self.ctx.raw, self.next_seed(), a.as_mut_ptr(), r.as_mut_ptr()) }, "qf_trap_gen");
        *self.installed.borrow_mut() = None; // qf_trap_gen installed a new key: whatever trapdoor was cached is gone
        let mut s = vec![0i64; m * m];
        self.ctx.check(unsafe { qf_gen_short_basis(self.ctx.raw, r.as_ptr(), s.as_mut_ptr()) }, "qf_gen_short_basis");
        let mut gso = vec![0f64; m * m];
        self.ctx.check(unsafe { qf_gso(self.ctx.raw, s.as_ptr(), gso.as_mut_ptr()) }, "qf_gso");
        let q = Z::from(&self.inner.gp.q);
        let (mut a_out, mut s_out, mut g_out) = (MatZq::new(n as i64, m as i64, &q), MatZ::new(m as i64, m as i64), MatQ::new(m as i64, m as i64));
        for i in 0..n {
            for j in 0..m {
                a_out.set_entry(i as i64, j as i64, Z::from(a[i * m + j])).unwrap();
            }
        }
        for i in 0..m {
            for j in 0..m {
                s_out.set_entry(i as i64, j as i64, Z::from(s[i * m + j])).unwrap();
                g_out.set_entry(i as i64, j as i64, Q::from(gso[i * m + j])).unwrap();
            }
        }
        (a_out, (s_out, g_out))
    }

    /// gpv.rs:113-116
    fn samp_d(&self) -> MatZ {
        let mut out = vec![0i32; self.m()];
        self.ctx.check(unsafe { qf_samp_d(self.ctx.raw, 1, self.next_seed(), 0, out.as_mut_ptr()) }, "qf_samp_d");
        column_from_i32(&out)
    }

    /// gpv.rs:152-161
    fn samp_p(&self, a: &MatZq, td: &(MatZ, MatQ), u: &MatZq) -> MatZ {
        self.samp_p_batch(a, td, std::slice::from_ref(u), self.next_seed(), 0).pop().unwrap()
    }

    /// gpv.rs:190-193: `assert!(self.check_domain(sigma)); a * sigma`
    fn f_a(&self, a: &MatZq, sigma: &MatZ) -> MatZq {
        self.f_a_batch(a, std::slice::from_ref(sigma)).pop().unwrap()
    }

    /// gpv.rs:219-224: column vector, length m, squared norm at most s^2 m
    fn check_domain(&self, sigma: &MatZ) -> bool {
        if !sigma.is_column_vector() || sigma.get_num_rows() as usize != self.m() {
            return false;
        }
        let Some(sg) = domain_to_i32(sigma) else { return false };
        let mut ok = [0u8; 1];
        self.ctx.check(unsafe { qf_check_domain(self.ctx.raw, sg.as_ptr(), 1, ok.as_mut_ptr()) }, "qf_check_domain");
        ok[0] == 1
    }
}

impl PSFBatch for PSFGPVB200 {
    fn samp_p_batch(&self, a: &MatZq, td: &(MatZ, MatQ), us: &[MatZq], seed: u64, first_index: u64) -> Vec<MatZ> {
        self.install(a, td);
        let (n, m) = (self.n(), self.m());
        let mut uw = Vec::with_capacity(us.len() * n);
        for u in us {
            assert!(u.is_column_vector() && u.get_num_rows() as usize == n, "target must be an n x 1 vector");
            uw.extend(matzq_to_words(u));
        }
        let mut e = vec![0i32; us.len() * m];
        let st = unsafe { qf_samp_p(self.ctx.raw, uw.as_ptr(), us.len() as i64, seed, first_index, e.as_mut_ptr()) };
        self.ctx.check(st, "qf_samp_p"); // the reference unwrap()s here (gpv.rs:155,160)
        e.chunks(m).map(column_from_i32).collect()
    }

    fn f_a_batch(&self, a: &MatZq, sigmas: &[MatZ]) -> Vec<MatZq> {
        let aw = matzq_to_words(a);
        let same = self.installed.borrow().as_ref().map(|(a0, _)| *a0 == aw).unwrap_or(false);
        if !same {
            self.ctx.check(unsafe { qf_set_a(self.ctx.raw, aw.as_ptr()) }, "qf_set_a");
            *self.installed.borrow_mut() = None;
        }
        let (n, m) = (self.n(), self.m());
        let mut sg = Vec::with_capacity(sigmas.len() * m);
        for sigma in sigmas {
            // the shape half of check_domain (gpv.rs:221-222) stays on the host
            assert!(sigma.is_column_vector() && sigma.get_num_rows() as usize == m, "sigma is not in D_n");
            sg.extend(domain_to_i32(sigma).expect("sigma is not in D_n"));
        }
        let mut u = vec![0i64; sigmas.len() * n];
        let mut ok = vec![0u8; sigmas.len()];
        let st = unsafe { qf_f_a(self.ctx.raw, sg.as_ptr(), sigmas.len() as i64, u.as_mut_ptr(), ok.as_mut_ptr()) };
        assert!(st != QF_ERR_NOT_IN_DOMAIN && ok.iter().all(|f| *f == 1), "sigma is not in D_n"); // gpv.rs:191
        self.ctx.check(st, "qf_f_a");
        let q = Z::from(&self.inner.gp.q);
        u.chunks(n).map(|row| range_from_words(row, &q)).collect()
    }
}

// ---- LossyCompressionFIPS203 -----------------------------------------------------------------------------------------

/// `Compress_d` of every coefficient word (lossy_compression_fips203.rs:101-111); panics for `d < 1` like `:91-94`.
pub fn compress_words(coeffs: &[i64], d: u32, q: u64) -> Vec<i64> {
    assert!(d >= 1, "Performing this function with d < 1 implies reducing mod 1");
    let mut out = vec![0i64; coeffs.len()];
    let st = unsafe { qf_compress_i64(coeffs.as_ptr(), out.as_mut_ptr(), coeffs.len(), q, d, 0, std::ptr::null_mut()) };
    assert!(st == QF_OK, "qf_compress_i64: status {st}");
    out
}

/// `Decompress_d` (lossy_compression_fips203.rs:159-169), coefficients written unreduced like the reference.
pub fn decompress_words(compressed: &[i64], d: u32, q: u64) -> Vec<i64> {
    assert!(d >= 1, "Performing this function with d < 1 implies reducing mod 1");
    let mut out = vec![0i64; compressed.len()];
    let st = unsafe { qf_decompress_i64(compressed.as_ptr(), out.as_mut_ptr(), compressed.len(), q, d, 0, std::ptr::null_mut()) };
    assert!(st == QF_OK, "qf_decompress_i64: status {st}");
    out
}

/// `ByteEncode_d(Compress_d(f))` for degree-256 polynomials with q < 2^16 (FIPS 203 Algorithm 5): the packed wire form.
pub fn compress_encode(coeffs: &[u16], d: u32, q: u32) -> Vec<u8> {
    assert!(d >= 1 && coeffs.len() % 256 == 0);
    let npoly = coeffs.len() / 256;
    let mut out = vec![0u8; npoly * 32 * d as usize];
    let st = unsafe { qf_compress_encode_u16(coeffs.as_ptr(), out.as_mut_ptr(), npoly, q, d, 1, 0, std::ptr::null_mut()) };
    assert!(st == QF_OK, "qf_compress_encode_u16: status {st}");
    out
}

/// `Decompress_d(ByteDecode_d(b))` (FIPS 203 Algorithm 6).
pub fn decode_decompress(packed: &[u8], d: u32, q: u32) -> Vec<u16> {
    assert!(d >= 1 && packed.len() % (32 * d as usize) == 0);
    let npoly = packed.len() / (32 * d as usize);
    let mut out = vec![0u16; npoly * 256];
    let st = unsafe { qf_decode_decompress_u16(packed.as_ptr(), out.as_mut_ptr(), npoly, q, d, 1, 0, std::ptr::null_mut()) };
    assert!(st == QF_OK, "qf_decode_decompress_u16: status {st}");
    out
}
