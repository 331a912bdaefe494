//! Reference-side binding of bellman's prover hot path to libb200zk.so.
//!
//! This is the source a maintainer of the reference (UrosTesic/zcash-gpu-thesis, a librustzcash fork) adds to
//! `bellman/src/` -- `multiexp()`, `EvaluationDomain` and `create_proof` keep their signatures and route their bodies
//! through the C ABI of `include/b200zk.h` (crate `b200zk-sys`).  It is NOT compiled in the build container of this
//! repository (no rustc / cargo there); the same ABI is exercised by the ctypes mirror (`zcash-gpu-thesis_b200/bellman.py`)
//! and the C++ mirror (`include/b200zk.hpp`), which the parity tests call.
//!
//! Reference lines replaced: bellman/src/multiexp.rs:19-68, 285-335; bellman/src/domain.rs:83-132;
//! bellman/src/groth16/prover.rs:249-364; bellman/src/groth16/mod.rs:252-382; bellman/src/groth16/verifier.rs:18-66.
extern crate b200zk_sys;
extern crate bit_vec;
extern crate byteorder;
extern crate futures;
extern crate pairing;

use b200zk_sys::*;
use std::io;
use std::marker::PhantomData;
use std::mem;
use std::os::raw::{c_int, c_void};
use std::ptr;
use std::sync::Arc;

use futures::{Async, Future, Poll};
use pairing::{CurveAffine, CurveProjective, Engine, PrimeField};

/// bellman::SynthesisError (bellman/src/lib.rs:171-188) as far as this path raises it
#[derive(Debug)]
pub enum SynthesisError {
    UnexpectedIdentity,
    PolynomialDegreeTooLarge,
    IoError(io::Error),
}

/// status -> SynthesisError
pub fn to_err(st: c_int) -> SynthesisError {
    match st {
        B200ZK_ERR_UNEXPECTED_IDENTITY => SynthesisError::UnexpectedIdentity,
        B200ZK_ERR_UNEXPECTED_EOF => SynthesisError::IoError(io::Error::new(io::ErrorKind::UnexpectedEof, "expected more bases from source")),
        B200ZK_ERR_DEGREE_TOO_LARGE => SynthesisError::PolynomialDegreeTooLarge,
        B200ZK_ERR_DECODE => SynthesisError::IoError(io::Error::new(io::ErrorKind::InvalidData, "group decoding error")),
        _ => SynthesisError::IoError(io::Error::new(io::ErrorKind::Other, "b200zk device error")),
    }
}

/// bellman::multicore::Worker (multicore.rs:13-49): here one GPU context (= one CUDA stream).  `Clone` shares the context; every
/// entry point of the library holds the context's lock, so a shared Worker is safe (calls are serialised).
#[derive(Clone)]
pub struct Worker { ctx: Arc<CtxHandle> }
struct CtxHandle(*mut b200zk_ctx);
unsafe impl Send for CtxHandle {}
unsafe impl Sync for CtxHandle {}
impl Drop for CtxHandle { fn drop(&mut self) { unsafe { b200zk_destroy(self.0) } } }
impl Worker {
    pub fn new() -> Worker { Worker::on_device(0) }
    pub fn on_device(device: c_int) -> Worker {
        let mut ctx = ptr::null_mut();
        let st = unsafe { b200zk_init(device, &mut ctx) };
        assert_eq!(st, B200ZK_OK, "b200zk_init failed: no usable CUDA device (there is no CPU fallback)");
        Worker { ctx: Arc::new(CtxHandle(ctx)) }
    }
    pub fn raw(&self) -> *mut b200zk_ctx { self.ctx.0 }
}

/// One process, several GPUs (b200zk_init_multi): what a single zcashd process needs for prover.rs:289-318.
pub struct MultiWorker { g: *mut b200zk_group }
unsafe impl Send for MultiWorker {}
unsafe impl Sync for MultiWorker {}
impl MultiWorker {
    pub fn new(devices: &[c_int]) -> MultiWorker {
        let mut g = ptr::null_mut();
        let st = unsafe { b200zk_init_multi(devices.as_ptr(), devices.len() as c_int, &mut g) };
        assert_eq!(st, B200ZK_OK);
        MultiWorker { g }
    }
    pub fn raw(&self) -> *mut b200zk_group { self.g }
}
impl Drop for MultiWorker { fn drop(&mut self) { unsafe { b200zk_group_destroy(self.g) } } }

pub trait GroupId { const GROUP_ID: c_int; }

/// A SourceBuilder whose bases already live in HBM (multiexp.rs:34-68).  `Parameters` builds one per query (h, l, a, b_g1, b_g2)
/// when the CRS is loaded (rustzcash.rs:230) and hands out `(handle, offset)` from ParameterSource::get_* exactly where it used
/// to hand out `(Arc<Vec<G>>, usize)` (groth16/mod.rs:456-481).
#[derive(Clone)]
pub struct DeviceBases<G: CurveAffine> { handle: Arc<BasesHandle>, pub offset: usize, _m: PhantomData<G> }
struct BasesHandle(*mut b200zk_bases);
unsafe impl Send for BasesHandle {}
unsafe impl Sync for BasesHandle {}
impl Drop for BasesHandle { fn drop(&mut self) { unsafe { b200zk_bases_free(self.0) } } }

impl<G: CurveAffine + GroupId> DeviceBases<G> {
    /// G1Affine { x: Fq, y: Fq, infinity: bool } is 104 bytes in memory (x, y Montgomery limbs first); the thesis relies on the
    /// same layout when it hands points to OpenCL (multiexp.rs:2876-2889).  The library reads the Vec in place through a stride.
    pub fn upload(worker: &Worker, v: &[G], precompute: bool) -> Result<Self, SynthesisError> {
        let stride = mem::size_of::<G>();
        let inf_off = stride - 8; // the bool sits after the two coordinates
        let p = v.as_ptr() as *const u8;
        let mut h = ptr::null_mut();
        let st = unsafe { b200zk_bases_upload(worker.raw(), G::GROUP_ID, p as *const c_void, v.len(), stride, p.add(inf_off), stride, &mut h) };
        if st != B200ZK_OK { return Err(to_err(st)); }
        if precompute {
            let st = unsafe { b200zk_bases_precompute(worker.raw(), h, 0) };
            if st != B200ZK_OK { unsafe { b200zk_bases_free(h) }; return Err(to_err(st)); }
        }
        Ok(DeviceBases { handle: Arc::new(BasesHandle(h)), offset: 0, _m: PhantomData })
    }
    /// the `(bases, offset)` a ParameterSource returns for the aux part of the A / B queries (groth16/mod.rs:456-481)
    pub fn at(&self, offset: usize) -> Self { DeviceBases { handle: self.handle.clone(), offset, _m: PhantomData } }
}

/// The future multiexp() returns (`Box<Future<Item = G::Projective, Error = SynthesisError>>`, multiexp.rs:285-295): the job is
/// submitted at construction (copy stream + compute stream), `poll` blocks in b200zk_job_wait.  The prover keeps eight of these
/// in flight (prover.rs:289-318, 339-354).
pub struct MultiexpFuture<G: CurveAffine> {
    job: *mut b200zk_job,
    _keep: (Arc<Vec<<<G::Engine as Engine>::Fr as PrimeField>::Repr>>, Option<Vec<u8>>),
}
unsafe impl<G: CurveAffine> Send for MultiexpFuture<G> {}
impl<G: CurveAffine> Future for MultiexpFuture<G> {
    type Item = <G as CurveAffine>::Projective;
    type Error = SynthesisError;
    fn poll(&mut self) -> Poll<Self::Item, Self::Error> {
        let mut out = <G as CurveAffine>::Projective::zero(); // G1 { x, y, z }: 18 u64, G2: 36 u64 (ec.rs:20-24)
        let job = mem::replace(&mut self.job, ptr::null_mut());
        let st = unsafe { b200zk_job_wait(job, &mut out as *mut _ as *mut u64) };
        if st == B200ZK_OK { Ok(Async::Ready(out)) } else { Err(to_err(st)) }
    }
}

/// bellman::multiexp::multiexp (multiexp.rs:285-335) with device-resident bases.  `density` = the DensityTracker's bit-vec
/// expanded to one byte per exponent (None = FullDensity); the query-size assertion of multiexp.rs:302-307 stays with the caller.
pub fn multiexp<G: CurveAffine + GroupId>(pool: &Worker, bases: DeviceBases<G>, density: Option<Vec<u8>>,
                                          exponents: Arc<Vec<<<G::Engine as Engine>::Fr as PrimeField>::Repr>>)
    -> Box<Future<Item = <G as CurveAffine>::Projective, Error = SynthesisError>>
{
    if let Some(ref d) = density { assert!(d.len() == exponents.len()); }
    let mut job = ptr::null_mut();
    let st = unsafe { b200zk_multiexp_async(pool.raw(), bases.handle.0, bases.offset, exponents.as_ptr() as *const u64, exponents.len(),
                                            density.as_ref().map_or(ptr::null(), |d| d.as_ptr()), &mut job) };
    if st != B200ZK_OK { return Box::new(futures::future::err(to_err(st))); }
    Box::new(MultiexpFuture::<G> { job, _keep: (exponents, density) })
}

/// EvaluationDomain<E, Scalar<E>> transforms (domain.rs:83-132): Scalar<E>(Fr(FrRepr([u64; 4]))) is 32 bytes of Montgomery limbs,
/// the Vec is passed as is and transformed in place.  kind: B200ZK_FFT / IFFT / COSET_FFT / ICOSET_FFT.
pub fn domain_ntt<S>(worker: &Worker, coeffs: &mut [S], exp: u32, kind: c_int) -> Result<(), SynthesisError> {
    debug_assert_eq!(mem::size_of::<S>(), 32);
    debug_assert_eq!(coeffs.len(), 1usize << exp);
    let st = unsafe { b200zk_ntt(worker.raw(), coeffs.as_mut_ptr() as *mut u64, exp, kind) };
    if st == B200ZK_OK { Ok(()) } else { Err(to_err(st)) }
}

/// The H block of create_proof (prover.rs:256-287) in one call: a, b, c evaluation vectors padded to m = 2^exp -> the m - 1
/// canonical coefficients the H multiexp consumes.
pub fn h_coefficients<S, R: Default + Clone>(worker: &Worker, a: &[S], b: &[S], c: &[S], exp: u32) -> Result<Vec<R>, SynthesisError> {
    let m = 1usize << exp;
    debug_assert!(a.len() == m && b.len() == m && c.len() == m);
    let mut out = vec![R::default(); m - 1];
    let st = unsafe { b200zk_h_poly(worker.raw(), a.as_ptr() as *const u64, b.as_ptr() as *const u64, c.as_ptr() as *const u64, exp, out.as_mut_ptr() as *mut u64) };
    if st == B200ZK_OK { Ok(out) } else { Err(to_err(st)) }
}

/// groth16::Parameters resident in HBM, straight from the wire format (Parameters::read, groth16/mod.rs:287-382).
pub struct DeviceParameters { crs: *mut b200zk_crs }
unsafe impl Send for DeviceParameters {}
unsafe impl Sync for DeviceParameters {}
impl Drop for DeviceParameters { fn drop(&mut self) { unsafe { b200zk_crs_free(self.crs) } } }
impl DeviceParameters {
    pub fn read(worker: &Worker, bytes: &[u8], checked: bool) -> Result<Self, SynthesisError> {
        let mut crs = ptr::null_mut();
        let st = unsafe { b200zk_parameters_read(worker.raw(), bytes.as_ptr(), bytes.len(), checked as c_int, &mut crs) };
        if st != B200ZK_OK { return Err(to_err(st)); }
        let st = unsafe { b200zk_crs_precompute(worker.raw(), crs, 0) };
        if st != B200ZK_OK { unsafe { b200zk_crs_free(crs) }; return Err(to_err(st)); }
        Ok(DeviceParameters { crs })
    }
    pub fn write(&self, worker: &Worker) -> Result<Vec<u8>, SynthesisError> {
        let mut out = vec![0u8; unsafe { b200zk_parameters_size(self.crs) }];
        let st = unsafe { b200zk_parameters_write(worker.raw(), self.crs, out.as_mut_ptr(), out.len()) };
        if st == B200ZK_OK { Ok(out) } else { Err(to_err(st)) }
    }
    pub fn raw(&self) -> *const b200zk_crs { self.crs }
}

/// What ProvingAssignment holds after synthesis (prover.rs:84-99), in the layout the ABI takes: evaluations as Montgomery
/// limbs, assignments as canonical FrRepr (prover.rs:290-291), the three density bit-vecs expanded to bytes.
pub struct SynthesizedAssignment<'a> {
    pub a: &'a [[u64; 4]], pub b: &'a [[u64; 4]], pub c: &'a [[u64; 4]],
    pub input_repr: &'a [[u64; 4]], pub aux_repr: &'a [[u64; 4]],
    pub a_aux_density: &'a [u8], pub b_input_density: &'a [u8], pub b_aux_density: &'a [u8],
}

/// groth16::create_proof after circuit synthesis (prover.rs:249-364) -> Proof::write bytes (groth16/mod.rs:43-53).  A run of
/// proofs over one circuit (what librustzcash_sapling_spend_proof is called for, once per shielded input, rustzcash.rs:1375)
/// is proved in lock-step groups: five batched multiexps per group.
pub fn create_proofs(worker: &Worker, params: &DeviceParameters, provers: &[SynthesizedAssignment], rs: &[([u64; 4], [u64; 4])])
    -> Result<Vec<[u8; 192]>, SynthesisError>
{
    assert_eq!(provers.len(), rs.len());
    if provers.is_empty() { return Ok(vec![]); }
    let rows: Vec<b200zk_prove_input> = provers.iter().zip(rs).map(|(p, &(ref r, ref s))| b200zk_prove_input {
        a: p.a.as_ptr() as *const u64, b: p.b.as_ptr() as *const u64, c: p.c.as_ptr() as *const u64,
        inputs: p.input_repr.as_ptr() as *const u64, aux: p.aux_repr.as_ptr() as *const u64,
        a_aux_density: p.a_aux_density.as_ptr(), b_input_density: p.b_input_density.as_ptr(), b_aux_density: p.b_aux_density.as_ptr(),
        r: r.as_ptr(), s: s.as_ptr() }).collect();
    let mut out = vec![[0u8; 192]; provers.len()];
    let st = unsafe { b200zk_groth16_prove_batch_bytes(worker.raw(), params.raw(), rows.as_ptr(), rows.len(), provers[0].a.len(),
                                                       provers[0].input_repr.len(), provers[0].aux_repr.len(), 0, out.as_mut_ptr() as *mut u8) };
    if st == B200ZK_OK { Ok(out) } else { Err(to_err(st)) } // UnexpectedIdentity covers the subversion check (prover.rs:320-324)
}
