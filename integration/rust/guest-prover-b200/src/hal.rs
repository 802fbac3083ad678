//! B200Hal: risc0_zkp::hal::Hal over libzkb200 (one C call per method).  SOURCE ONLY -- there is no Rust toolchain in the
//! image this repository is built in, so this file documents the binding and is not compiled.
//! The Python mirror that IS exercised by the parity tests is zktls_b200/hal.py; method for method they are the same.
use std::{cell::RefCell, ffi::c_void, marker::PhantomData, rc::Rc};

use risc0_core::field::baby_bear::{BabyBear, BabyBearElem, BabyBearExtElem};
use risc0_zkp::{core::{digest::Digest, hash::{poseidon2::Poseidon2HashSuite, HashSuite}, log2_ceil}, hal::{Buffer, Hal}};
use zkb200_sys as sys;

pub(crate) fn ok(e: sys::ZkbErr) { sys::ffi_wrap(|| e).unwrap_or_else(|m| panic!("zkb200: {m}")) }

struct Raw { ctx: *mut sys::ZkbCtx, ptr: *mut c_void, bytes: usize }
impl Drop for Raw { fn drop(&mut self) { unsafe { sys::zkb_free(self.ctx, self.ptr); } } }

/// Cheap-clone device buffer handle (the reference's CudaHal buffer is an Rc<RefCell<..>> too, hence !Send).
#[derive(Clone)]
pub struct B200Buffer<T> { raw: Rc<RefCell<Raw>>, offset: usize, size: usize, _t: PhantomData<T> }

impl<T: Clone + bytemuck::Pod> B200Buffer<T> {
    pub fn as_device_ptr(&self) -> *mut c_void { unsafe { (self.raw.borrow().ptr as *mut u8).add(self.offset * std::mem::size_of::<T>()) as *mut c_void } }
}
impl<T: Clone + bytemuck::Pod> Buffer<T> for B200Buffer<T> {
    fn name(&self) -> &'static str { "b200" }
    fn size(&self) -> usize { self.size }
    fn slice(&self, offset: usize, size: usize) -> Self { assert!(offset + size <= self.size); Self { raw: self.raw.clone(), offset: self.offset + offset, size, _t: PhantomData } }
    fn get_at(&self, idx: usize) -> T { let mut v = [T::zeroed()]; self.view_into(idx, &mut v); v[0] }
    fn view<F: FnOnce(&[T])>(&self, f: F) { let mut v = vec![T::zeroed(); self.size]; self.view_into(0, &mut v); f(&v) }
    fn view_mut<F: FnOnce(&mut [T])>(&self, f: F) {
        let mut v = vec![T::zeroed(); self.size]; self.view_into(0, &mut v); f(&mut v);
        let r = self.raw.borrow();
        ok(unsafe { sys::zkb_h2d(r.ctx, self.as_device_ptr(), v.as_ptr() as *const c_void, v.len() * std::mem::size_of::<T>()) });
    }
    fn to_vec(&self) -> Vec<T> { let mut v = vec![T::zeroed(); self.size]; self.view_into(0, &mut v); v }
}
impl<T: Clone + bytemuck::Pod> B200Buffer<T> {
    fn view_into(&self, idx: usize, out: &mut [T]) {
        let r = self.raw.borrow();
        let src = unsafe { (self.as_device_ptr() as *const u8).add(idx * std::mem::size_of::<T>()) } as *const c_void;
        ok(unsafe { sys::zkb_d2h(r.ctx, out.as_mut_ptr() as *mut c_void, src, out.len() * std::mem::size_of::<T>()) });   // blocks, like CudaHal
    }
}

/// `suite`: the `poseidon2` hash suite (host-side hashing of small inputs + the Fiat-Shamir rng; the device kernels implement the same
/// permutation -- libzkb200 has no other suite, which is why there is no `B200Hal<Hash>` type parameter like `CudaHal<CH>`).
pub struct B200Hal { pub(crate) ctx: *mut sys::ZkbCtx, suite: HashSuite<BabyBear> }

impl B200Hal {
    pub fn new(device: i32) -> Self { let mut ctx = std::ptr::null_mut(); ok(unsafe { sys::zkb_init(device, &mut ctx) }); Self { ctx, suite: Poseidon2HashSuite::new_suite() } }
    fn alloc<T: Clone + bytemuck::Pod>(&self, size: usize) -> B200Buffer<T> {
        let bytes = (size * std::mem::size_of::<T>()).max(16); let mut p = std::ptr::null_mut();
        ok(unsafe { sys::zkb_alloc(self.ctx, bytes, &mut p) }); ok(unsafe { sys::zkb_memset0(self.ctx, p, bytes) });
        B200Buffer { raw: Rc::new(RefCell::new(Raw { ctx: self.ctx, ptr: p, bytes })), offset: 0, size, _t: PhantomData }
    }
    fn upload<T: Clone + bytemuck::Pod>(&self, s: &[T]) -> B200Buffer<T> {
        let b = self.alloc::<T>(s.len());
        ok(unsafe { sys::zkb_h2d(self.ctx, b.as_device_ptr(), s.as_ptr() as *const c_void, std::mem::size_of_val(s)) }); b
    }
}
impl Drop for B200Hal { fn drop(&mut self) { unsafe { sys::zkb_destroy(self.ctx); } } }

impl Hal for B200Hal {
    type Field = BabyBear; type Elem = BabyBearElem; type ExtElem = BabyBearExtElem; type Buffer<T: Clone + bytemuck::Pod> = B200Buffer<T>;

    fn has_unified_memory(&self) -> bool { false }
    fn get_hash_suite(&self) -> &HashSuite<BabyBear> { &self.suite }

    fn alloc_elem(&self, _n: &'static str, size: usize) -> Self::Buffer<Self::Elem> { self.alloc(size) }
    fn alloc_extelem(&self, _n: &'static str, size: usize) -> Self::Buffer<Self::ExtElem> { self.alloc(size) }
    fn alloc_digest(&self, _n: &'static str, size: usize) -> Self::Buffer<Digest> { self.alloc(size) }
    fn alloc_u32(&self, _n: &'static str, size: usize) -> Self::Buffer<u32> { self.alloc(size) }
    fn copy_from_elem(&self, _n: &'static str, s: &[Self::Elem]) -> Self::Buffer<Self::Elem> { self.upload(s) }
    fn copy_from_extelem(&self, _n: &'static str, s: &[Self::ExtElem]) -> Self::Buffer<Self::ExtElem> { self.upload(s) }
    fn copy_from_digest(&self, _n: &'static str, s: &[Digest]) -> Self::Buffer<Digest> { self.upload(s) }
    fn copy_from_u32(&self, _n: &'static str, s: &[u32]) -> Self::Buffer<u32> { self.upload(s) }

    fn batch_interpolate_ntt(&self, io: &Self::Buffer<Self::Elem>, count: usize) {
        ok(unsafe { sys::zkb_batch_interpolate_ntt(self.ctx, io.as_device_ptr(), count, log2_ceil(io.size() / count) as i32) })
    }
    fn zk_shift(&self, io: &Self::Buffer<Self::Elem>, count: usize) {
        ok(unsafe { sys::zkb_zk_shift(self.ctx, io.as_device_ptr(), count, log2_ceil(io.size() / count) as i32) })
    }
    fn batch_expand_into_evaluate_ntt(&self, out: &Self::Buffer<Self::Elem>, inp: &Self::Buffer<Self::Elem>, count: usize, expand_bits: usize) {
        ok(unsafe { sys::zkb_batch_expand_into_evaluate_ntt(self.ctx, out.as_device_ptr(), inp.as_device_ptr(), count, log2_ceil(inp.size() / count) as i32, expand_bits as i32) })
    }
    fn batch_bit_reverse(&self, io: &Self::Buffer<Self::Elem>, count: usize) {
        ok(unsafe { sys::zkb_batch_bit_reverse(self.ctx, io.as_device_ptr(), count, log2_ceil(io.size() / count) as i32) })
    }
    fn hash_rows(&self, output: &Self::Buffer<Digest>, matrix: &Self::Buffer<Self::Elem>) {
        let rows = output.size();
        ok(unsafe { sys::zkb_poseidon2_hash_rows(self.ctx, output.as_device_ptr(), matrix.as_device_ptr(), rows, matrix.size() / rows) })
    }
    fn hash_fold(&self, io: &Self::Buffer<Digest>, input_size: usize, output_size: usize) {
        ok(unsafe { sys::zkb_poseidon2_hash_fold(self.ctx, io.as_device_ptr(), input_size, output_size) })
    }
    fn batch_evaluate_any(&self, coeffs: &Self::Buffer<Self::Elem>, poly_count: usize, which: &Self::Buffer<u32>, xs: &Self::Buffer<Self::ExtElem>, out: &Self::Buffer<Self::ExtElem>) {
        ok(unsafe { sys::zkb_batch_evaluate_any(self.ctx, coeffs.as_device_ptr(), poly_count, log2_ceil(coeffs.size() / poly_count) as i32, which.as_device_ptr(), xs.as_device_ptr(), out.as_device_ptr(), which.size()) })
    }
    fn mix_poly_coeffs(&self, out: &Self::Buffer<Self::ExtElem>, mix_start: &Self::ExtElem, mix: &Self::ExtElem, input: &Self::Buffer<Self::Elem>, combos: &Self::Buffer<u32>, input_size: usize, count: usize) {
        ok(unsafe { sys::zkb_mix_poly_coeffs(self.ctx, out.as_device_ptr(), mix_start as *const _ as *const u32, mix as *const _ as *const u32, input.as_device_ptr(), combos.as_device_ptr(), input_size, count) })
    }
    fn eltwise_sum_extelem(&self, out: &Self::Buffer<Self::Elem>, input: &Self::Buffer<Self::ExtElem>) {
        let count = out.size() / 4;
        ok(unsafe { sys::zkb_eltwise_sum_extelem(self.ctx, out.as_device_ptr(), input.as_device_ptr(), count, input.size() / count) })
    }
    fn fri_fold(&self, out: &Self::Buffer<Self::Elem>, input: &Self::Buffer<Self::Elem>, mix: &Self::ExtElem) {
        ok(unsafe { sys::zkb_fri_fold(self.ctx, out.as_device_ptr(), input.as_device_ptr(), mix as *const _ as *const u32, out.size() / 4) })
    }
    fn eltwise_add_elem(&self, out: &Self::Buffer<Self::Elem>, a: &Self::Buffer<Self::Elem>, b: &Self::Buffer<Self::Elem>) {
        ok(unsafe { sys::zkb_eltwise_add_elem(self.ctx, out.as_device_ptr(), a.as_device_ptr(), b.as_device_ptr(), out.size()) })
    }
    fn eltwise_copy_elem(&self, out: &Self::Buffer<Self::Elem>, input: &Self::Buffer<Self::Elem>) {
        ok(unsafe { sys::zkb_eltwise_copy_elem(self.ctx, out.as_device_ptr(), input.as_device_ptr(), out.size()) })
    }
    fn eltwise_zeroize_elem(&self, io: &Self::Buffer<Self::Elem>) { ok(unsafe { sys::zkb_eltwise_zeroize_elem(self.ctx, io.as_device_ptr(), io.size()) }) }
    fn gather_sample(&self, dst: &Self::Buffer<Self::Elem>, src: &Self::Buffer<Self::Elem>, idx: usize, size: usize, stride: usize) {
        ok(unsafe { sys::zkb_gather_sample(self.ctx, dst.as_device_ptr(), src.as_device_ptr(), idx, size, stride) })
    }
    fn prefix_products(&self, io: &Self::Buffer<Self::ExtElem>) { ok(unsafe { sys::zkb_prefix_products(self.ctx, io.as_device_ptr(), io.size()) }) }
    fn scatter(&self, into: &Self::Buffer<Self::Elem>, index: &[u32], offsets: &[u32], values: &[Self::Elem]) {
        ok(unsafe { sys::zkb_scatter(self.ctx, into.as_device_ptr(), into.size(), index.as_ptr(), index.len() - 1, offsets.as_ptr(), values.as_ptr() as *const u32) })
    }
}

/// B200CircuitHal: risc0_zkp::hal::CircuitHal<B200Hal> for a circuit described by a blob (TapSet + PolyExtStep program).
/// SOURCE ONLY, like the rest of this file.  `blob` is built once per circuit from `CircuitDef::{taps, poly_ext}`.
pub struct B200CircuitHal { ctx: *mut sys::ZkbCtx, blob: Vec<u32> }

impl B200CircuitHal {
    pub fn new(hal: &B200Hal, blob: Vec<u32>) -> Self {
        ok(unsafe { sys::zkb_eval_check_precompile(blob.as_ptr(), blob.len()) });     // NVRTC-specialise eval_check ahead of the first proof
        Self { ctx: hal.ctx, blob }
    }
}

impl risc0_zkp::hal::CircuitHal<B200Hal> for B200CircuitHal {
    /// groups = [accum, code, data] LDE matrices; globals = [mix, out] host-visible (small) buffers
    fn eval_check(&self, check: &B200Buffer<BabyBearElem>, groups: &[&B200Buffer<BabyBearElem>], globals: &[&B200Buffer<BabyBearElem>],
                  poly_mix: BabyBearExtElem, po2: usize, _steps: usize) {
        let (mix, out) = (globals[0].to_vec(), globals[1].to_vec());
        ok(unsafe { sys::zkb_eval_check(self.ctx, check.as_device_ptr(), self.blob.as_ptr(), self.blob.len(), groups[0].as_device_ptr(), groups[1].as_device_ptr(),
                                        groups[2].as_device_ptr(), mix.as_ptr() as *const u32, out.as_ptr() as *const u32, &poly_mix as *const _ as *const u32, po2 as i32) })
    }
    /// ctrl = the code trace.  The witness program is the circuit's: libzkb200 carries the SYN family's; the rv32im step functions
    /// (`step_compute_accum` / `step_verify_accum`) would be added to csrc/k_accum.cu the same way eval_check takes `poly_ext` as data.
    fn accumulate(&self, ctrl: &B200Buffer<BabyBearElem>, io: &B200Buffer<BabyBearElem>, data: &B200Buffer<BabyBearElem>, mix: &B200Buffer<BabyBearElem>,
                  accum: &B200Buffer<BabyBearElem>, steps: usize) {
        let (mix, io) = (mix.to_vec(), io.to_vec());
        ok(unsafe { sys::zkb_accumulate(self.ctx, self.blob.as_ptr(), self.blob.len(), accum.as_device_ptr(), ctrl.as_device_ptr(), data.as_device_ptr(),
                                        mix.as_ptr() as *const u32, io.as_ptr() as *const u32, log2_ceil(steps) as i32) })
    }
}
