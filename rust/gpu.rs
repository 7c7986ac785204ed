//! GPU voice banks for FunDSP graphs: binding of `libfundsp_b200.so` (C ABI: include/fundsp_b200.h).
//!
//! Written against fundsp v0.23.0. NOT compiled in the fundsp_b200 repository (no Rust toolchain in its build image); the C++ and
//! Python mirrors of this file are what its tests run. Lives in the crate (`src/gpu.rs`, feature `gpu`) because lowering reads
//! construction-time parameters that are private fields; it needs these `pub(crate)` accessors added next to the fields:
//!   Sine::initial_phase() -> Option<f32>            (oscillator.rs:25)     WaveSynth::initial_phase() / table_kind() (wavetable.rs:258)
//!   Noise::seed() -> Option<u64>                    (noise.rs:175)         FixedSvf::params() -> &SvfParams<f32>     (svf.rs:867)
//!   Moog::cutoff_q() -> (f32, f32)                  (moog.rs:20-34)        Fir::weights() -> &[f32]                  (fir.rs:15)
//!   Delay::length_seconds() -> f64                  (delay.rs:69)          Panner::<U1>::pan_value() -> f32          (pan.rs:19)
//!   Unop scalar: FrameAddScalar / FrameMulScalar / FrameNegAddScalar ::scalar() (audionode.rs:1114,1155,1197)
//! Everything else goes through the public API (`Pipe::left()/right()`, `Constant::value()`, `AudioNode::ID`, ...).
#![allow(clippy::missing_safety_doc)]
use crate::audionode::*;
use crate::audiounit::AudioUnit;
use crate::buffer::{BufferMut, BufferRef};
use crate::combinator::An;
use crate::math::AttoHash;
use crate::setting::{Address, Parameter, Setting};
use crate::signal::{Routing, SignalFrame};
use crate::*;
use core::ffi::{c_char, c_int, c_void};
use numeric_array::typenum::*;
extern crate alloc;
use alloc::{string::String, vec::Vec};

#[repr(C)] pub struct FdspNode { _p: [u8; 0] }
#[repr(C)] pub struct FdspBank { _p: [u8; 0] }
#[repr(C)] pub struct FdspGroup { _p: [u8; 0] }

pub const FDSP_OUT_VOICES: u32 = 1;
pub const FDSP_OUT_MIX: u32 = 2;

#[link(name = "fundsp_b200")]
extern "C" {
    fn fdsp_last_error() -> *const c_char;
    fn fdsp_constant(n: c_int, values: *const f32) -> *mut FdspNode;
    fn fdsp_pass() -> *mut FdspNode;
    fn fdsp_multipass(n: c_int) -> *mut FdspNode;
    fn fdsp_sink(n: c_int) -> *mut FdspNode;
    fn fdsp_multisplit(m: c_int, n: c_int) -> *mut FdspNode;
    fn fdsp_multijoin(m: c_int, n: c_int) -> *mut FdspNode;
    fn fdsp_sine() -> *mut FdspNode;
    fn fdsp_wavesynth(table: c_int, outputs: c_int) -> *mut FdspNode;
    fn fdsp_noise() -> *mut FdspNode;
    fn fdsp_fixed_svf(mode: c_int, cutoff: f32, q: f32, gain: f32) -> *mut FdspNode;
    fn fdsp_moog(cutoff: f32, q: f32, inputs: c_int) -> *mut FdspNode;
    fn fdsp_fir(n: c_int, weights: *const f32) -> *mut FdspNode;
    fn fdsp_delay(seconds: f64) -> *mut FdspNode;
    fn fdsp_pan(value: f32) -> *mut FdspNode;
    fn fdsp_adsr_live(a: f32, d: f32, s: f32, r: f32) -> *mut FdspNode;
    fn fdsp_convolve(response: *const f32, n: c_int) -> *mut FdspNode;
    fn fdsp_pipe(x: *mut FdspNode, y: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_stack(x: *mut FdspNode, y: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_branch(x: *mut FdspNode, y: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_bus(x: *mut FdspNode, y: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_thru(x: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_binop(op: c_int, x: *mut FdspNode, y: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_unop(kind: c_int, scalar: f32, x: *mut FdspNode) -> *mut FdspNode;
    fn fdsp_multi(kind: c_int, op: c_int, n: c_int, nodes: *const *mut FdspNode) -> *mut FdspNode;
    fn fdsp_feedback(x: *mut FdspNode, hadamard: c_int) -> *mut FdspNode;
    fn fdsp_node_phase(n: *mut FdspNode, phase: f32) -> c_int;
    fn fdsp_node_seed(n: *mut FdspNode, seed: u64) -> c_int;
    fn fdsp_node_free(n: *mut FdspNode);
    fn fdsp_bank_create(voices: *const *mut FdspNode, n: u32, device: c_int, out_mode: u32, out: *mut *mut FdspBank) -> c_int;
    fn fdsp_bank_destroy(b: *mut FdspBank);
    fn fdsp_bank_clone(b: *const FdspBank, out: *mut *mut FdspBank) -> c_int;
    fn fdsp_bank_inputs(b: *const FdspBank) -> c_int;
    fn fdsp_bank_outputs(b: *const FdspBank) -> c_int;
    fn fdsp_bank_set_sample_rate(b: *mut FdspBank, sr: f64) -> c_int;
    fn fdsp_bank_reset(b: *mut FdspBank) -> c_int;
    fn fdsp_bank_allocate(b: *mut FdspBank, max_render_samples: u64) -> c_int;
    fn fdsp_bank_set(b: *mut FdspBank, voice: u32, kind: c_int, v: *const f32, nv: c_int, seed: u64, addr: *const i64, naddr: c_int) -> c_int;
    fn fdsp_bank_process(b: *mut FdspBank, size: u32, input: *const f32, output: *mut f32) -> c_int;
    fn fdsp_bank_render(b: *mut FdspBank, n: u64, input: *const f32, out_voices: *mut f32, out_mix: *mut f32) -> c_int;
    // multi-GPU mix-down (csrc/host/group.h)
    fn fdsp_group_unique_id(id: *mut c_void, bytes: u64) -> c_int;
    fn fdsp_group_create(nranks: c_int, rank: c_int, id: *const c_void, device: c_int, out: *mut *mut FdspGroup) -> c_int;
    fn fdsp_group_destroy(g: *mut FdspGroup);
    fn fdsp_bank_render_reduced(b: *mut FdspBank, g: *mut FdspGroup, n: u64, input: *const f32, out_mix: *mut f32, root: c_int) -> c_int;
}

fn last_error() -> String {
    unsafe {
        let p = fdsp_last_error();
        if p.is_null() { return String::new(); }
        let mut n = 0usize;
        while *p.add(n) != 0 { n += 1; }
        String::from_utf8_lossy(core::slice::from_raw_parts(p as *const u8, n)).into_owned()
    }
}
fn check(rc: c_int) -> Result<(), String> { if rc == 0 { Ok(()) } else { Err(last_error()) } }

/// A typed graph lowers itself: one builder call per node, depth first, left to right — the order the reference constructs and pings
/// its nodes in (audionode.rs:156-161), so the deterministic phase hashes of the GPU voices equal the ones the Rust tree computed.
pub trait Lower { unsafe fn lower(&self) -> *mut FdspNode; }

impl<X: AudioNode + Lower> Lower for An<X> { unsafe fn lower(&self) -> *mut FdspNode { self.0.lower() } }

impl<X, Y> Lower for Pipe<X, Y> where X: AudioNode + Lower, Y: AudioNode<Inputs = X::Outputs> + Lower {
    unsafe fn lower(&self) -> *mut FdspNode { fdsp_pipe(self.left().lower(), self.right().lower()) }            // audionode.rs:1370
}
impl<X, Y> Lower for Stack<X, Y> where X: AudioNode + Lower, Y: AudioNode + Lower, X::Inputs: core::ops::Add<Y::Inputs>, X::Outputs: core::ops::Add<Y::Outputs>,
    <X::Inputs as core::ops::Add<Y::Inputs>>::Output: Size<f32>, <X::Outputs as core::ops::Add<Y::Outputs>>::Output: Size<f32> {
    unsafe fn lower(&self) -> *mut FdspNode { fdsp_stack(self.left().lower(), self.right().lower()) }           // audionode.rs:1494
}
impl<X, Y> Lower for Branch<X, Y> where X: AudioNode + Lower, Y: AudioNode<Inputs = X::Inputs> + Lower, X::Outputs: core::ops::Add<Y::Outputs>,
    <X::Outputs as core::ops::Add<Y::Outputs>>::Output: Size<f32> {
    unsafe fn lower(&self) -> *mut FdspNode { fdsp_branch(self.left().lower(), self.right().lower()) }          // audionode.rs:1651
}
impl<X, Y> Lower for Bus<X, Y> where X: AudioNode + Lower, Y: AudioNode<Inputs = X::Inputs, Outputs = X::Outputs> + Lower {
    unsafe fn lower(&self) -> *mut FdspNode { fdsp_bus(self.left().lower(), self.right().lower()) }             // audionode.rs:1794
}
impl<X: AudioNode + Lower> Lower for Thru<X> { unsafe fn lower(&self) -> *mut FdspNode { fdsp_thru(self.inner().lower()) } }   // audionode.rs:1948

/// `+ - *` between graphs: Binop<FrameAdd|FrameSub|FrameMul, X, Y> (audionode.rs:850-1027); op 0 add, 1 sub, 2 mul.
pub trait BinopCode { const OP: c_int; }
impl<N: Size<f32>> BinopCode for FrameAdd<N> { const OP: c_int = 0; }
impl<N: Size<f32>> BinopCode for FrameSub<N> { const OP: c_int = 1; }
impl<N: Size<f32>> BinopCode for FrameMul<N> { const OP: c_int = 2; }
impl<B, X, Y> Lower for Binop<B, X, Y> where B: FrameBinop<X::Outputs> + BinopCode, X: AudioNode + Lower, Y: AudioNode<Outputs = X::Outputs> + Lower,
    X::Inputs: core::ops::Add<Y::Inputs>, <X::Inputs as core::ops::Add<Y::Inputs>>::Output: Size<f32> {
    unsafe fn lower(&self) -> *mut FdspNode { fdsp_binop(B::OP, self.left().lower(), self.right().lower()) }
}
/// `-x`, `x + s`, `s - x`, `x * s`: Unop (audionode.rs:1229-1326); kind 0 neg, 1 +s, 2 -x+s, 3 *s. Needs the `scalar()` accessors.
pub trait UnopCode { const KIND: c_int; fn scalar_value(&self) -> f32; }
impl<N: Size<f32>> UnopCode for FrameNeg<N> { const KIND: c_int = 0; fn scalar_value(&self) -> f32 { 0.0 } }
impl<N: Size<f32>> UnopCode for FrameAddScalar<N> { const KIND: c_int = 1; fn scalar_value(&self) -> f32 { self.scalar() } }
impl<N: Size<f32>> UnopCode for FrameNegAddScalar<N> { const KIND: c_int = 2; fn scalar_value(&self) -> f32 { self.scalar() } }
impl<N: Size<f32>> UnopCode for FrameMulScalar<N> { const KIND: c_int = 3; fn scalar_value(&self) -> f32 { self.scalar() } }
impl<X, U> Lower for Unop<X, U> where X: AudioNode + Lower, U: FrameUnop<X::Outputs> + UnopCode {
    unsafe fn lower(&self) -> *mut FdspNode { fdsp_unop(U::KIND, self.unop().scalar_value(), self.inner().lower()) }
}

impl<N: Size<f32>> Lower for Constant<N> {
    unsafe fn lower(&self) -> *mut FdspNode { let v = self.value(); fdsp_constant(N::I32, v.as_ptr()) }        // audionode.rs:465
}
impl Lower for Pass { unsafe fn lower(&self) -> *mut FdspNode { fdsp_pass() } }
impl<N: Size<f32>> Lower for MultiPass<N> { unsafe fn lower(&self) -> *mut FdspNode { fdsp_multipass(N::I32) } }
impl<N: Size<f32>> Lower for Sink<N> { unsafe fn lower(&self) -> *mut FdspNode { fdsp_sink(N::I32) } }
impl Lower for crate::oscillator::Sine<f32> {
    unsafe fn lower(&self) -> *mut FdspNode { let n = fdsp_sine(); if let Some(p) = self.initial_phase() { fdsp_node_phase(n, p); } n }   // `.phase(p)`
}
impl<N: Size<f32>> Lower for crate::wavetable::WaveSynth<N> {
    unsafe fn lower(&self) -> *mut FdspNode {
        // table_kind(): 0 saw, 1 square, 2 triangle, 3 organ, 4 soft saw, 5 hammond (the lazily built global tables of wavetable.rs:493-623)
        let n = fdsp_wavesynth(self.table_kind() as c_int, N::I32);
        if let Some(p) = self.initial_phase() { fdsp_node_phase(n, p); }
        n
    }
}
impl Lower for crate::noise::Noise {
    unsafe fn lower(&self) -> *mut FdspNode { let n = fdsp_noise(); if let Some(s) = self.seed() { fdsp_node_seed(n, s); } n }            // `.seed(s)`
}
/// SVF mode index: 0 lowpass 1 highpass 2 bandpass 3 notch 4 peak 5 allpass 6 bell 7 lowshelf 8 highshelf (svf.rs:26-221).
pub trait SvfModeIndex { const INDEX: c_int; }
impl SvfModeIndex for crate::svf::LowpassMode<f32> { const INDEX: c_int = 0; }
impl SvfModeIndex for crate::svf::HighpassMode<f32> { const INDEX: c_int = 1; }
impl SvfModeIndex for crate::svf::BandpassMode<f32> { const INDEX: c_int = 2; }
impl SvfModeIndex for crate::svf::NotchMode<f32> { const INDEX: c_int = 3; }
impl SvfModeIndex for crate::svf::PeakMode<f32> { const INDEX: c_int = 4; }
impl SvfModeIndex for crate::svf::AllpassMode<f32> { const INDEX: c_int = 5; }
impl SvfModeIndex for crate::svf::BellMode<f32> { const INDEX: c_int = 6; }
impl SvfModeIndex for crate::svf::LowshelfMode<f32> { const INDEX: c_int = 7; }
impl SvfModeIndex for crate::svf::HighshelfMode<f32> { const INDEX: c_int = 8; }
impl<M: crate::svf::SvfMode<f32> + SvfModeIndex> Lower for crate::svf::FixedSvf<f32, M> {
    unsafe fn lower(&self) -> *mut FdspNode { let p = self.params(); fdsp_fixed_svf(M::INDEX, p.cutoff, p.q, p.gain) }                   // svf.rs:857
}
impl<N: Size<f32>> Lower for crate::moog::Moog<f32, N> {
    unsafe fn lower(&self) -> *mut FdspNode { let (c, q) = self.cutoff_q(); fdsp_moog(c, q, N::I32) }                                    // moog.rs:11-46
}
impl<N: Size<f32>> Lower for crate::fir::Fir<N> {
    unsafe fn lower(&self) -> *mut FdspNode { let w = self.weights(); fdsp_fir(N::I32, w.as_ptr()) }                                     // fir.rs:11-41
}
impl Lower for crate::delay::Delay { unsafe fn lower(&self) -> *mut FdspNode { fdsp_delay(self.length_seconds()) } }                       // delay.rs:69
impl Lower for crate::pan::Panner<U1> { unsafe fn lower(&self) -> *mut FdspNode { fdsp_pan(self.pan_value()) } }                           // pan.rs:19
impl Lower for crate::convolve::Convolver {
    unsafe fn lower(&self) -> *mut FdspNode { let h = self.response(); fdsp_convolve(h.as_ptr(), h.len() as c_int) }                     // convolve.rs:9
}
/// `adsr_live(a, d, s, r)` is `EnvelopeIn` with the closure of adsr.rs:21-70: closures do not cross a C ABI, the closed form does.
pub struct AdsrLive { pub attack: f32, pub decay: f32, pub sustain: f32, pub release: f32 }
impl Lower for AdsrLive { unsafe fn lower(&self) -> *mut FdspNode { fdsp_adsr_live(self.attack, self.decay, self.sustain, self.release) } }

/// V voices of typed graphs evaluated in lockstep on one GPU; to the host ONE `AudioUnit` (audiounit.rs:21-95).
pub struct GpuBank { h: *mut FdspBank, inputs: usize, outputs: usize, failed: bool }
unsafe impl Send for GpuBank {}
unsafe impl Sync for GpuBank {}

impl GpuBank {
    /// `voices`: per-voice graphs (they may fall into several structural classes); `mix`: outputs() = channels of the summed voices,
    /// else V * channels per-voice outputs.
    pub fn new<X: AudioNode + Lower>(voices: &[An<X>], device: i32, mix: bool) -> Result<Self, String> {
        let mut hs: Vec<*mut FdspNode> = Vec::with_capacity(voices.len());
        for v in voices {
            let h = unsafe { v.lower() };
            if h.is_null() { for x in hs { unsafe { fdsp_node_free(x) } } return Err(last_error()); }
            hs.push(h);
        }
        let mut b: *mut FdspBank = core::ptr::null_mut();
        check(unsafe { fdsp_bank_create(hs.as_ptr(), hs.len() as u32, device, if mix { FDSP_OUT_MIX } else { FDSP_OUT_VOICES }, &mut b) })?;   // consumes the handles
        let (i, o) = unsafe { (fdsp_bank_inputs(b) as usize, fdsp_bank_outputs(b) as usize) };
        Ok(GpuBank { h: b, inputs: i, outputs: o, failed: false })
    }
    /// `Wave::render` for many blocks in one call (wave.rs:441-466): `out` is [outputs()][samples], channel-major.
    pub fn render(&mut self, samples: usize, input: Option<&[f32]>, out: &mut [f32]) -> Result<(), String> {
        assert!(out.len() >= self.outputs * samples);
        check(unsafe { fdsp_bank_render(self.h, samples as u64, input.map_or(core::ptr::null(), |x| x.as_ptr()), core::ptr::null_mut(), out.as_mut_ptr()) })
    }
    pub fn failed(&self) -> bool { self.failed }
}
impl AudioUnit for GpuBank {
    fn inputs(&self) -> usize { self.inputs }
    fn outputs(&self) -> usize { self.outputs }
    fn reset(&mut self) { if unsafe { fdsp_bank_reset(self.h) } != 0 { self.failed = true; } }
    fn set_sample_rate(&mut self, sr: f64) { if unsafe { fdsp_bank_set_sample_rate(self.h, sr) } != 0 { self.failed = true; } }
    fn allocate(&mut self) { unsafe { fdsp_bank_allocate(self.h, 64); } }                    // later process() calls do not allocate (audiounit.rs:92-95)
    fn set(&mut self, setting: Setting) {                                                     // audiounit.rs:62: address[0] = Index(voice)
        let mut s = setting;
        let voice = match s.direction() { Address::Index(i) => i as u32, _ => return };        // silently ignored like net.rs:1166
        s = s.peel();
        let (kind, vals): (c_int, Vec<f32>) = match s.parameter() {
            Parameter::Center(c) => (1, alloc::vec![*c]), Parameter::CenterQ(c, q) => (2, alloc::vec![*c, *q]), Parameter::CenterQGain(c, q, g) => (3, alloc::vec![*c, *q, *g]),
            Parameter::Value(v) => (4, alloc::vec![*v]), Parameter::Coefficient(c) => (5, alloc::vec![*c]),
            Parameter::Biquad(a1, a2, b0, b1, b2) => (6, alloc::vec![*a1, *a2, *b0, *b1, *b2]), Parameter::Delay(d) => (7, alloc::vec![*d]), Parameter::Time(t) => (8, alloc::vec![*t]),
            Parameter::Roughness(r) => (9, alloc::vec![*r]), Parameter::Variability(v) => (10, alloc::vec![*v]), Parameter::Pan(p) => (11, alloc::vec![*p]),
            Parameter::AttackRelease(a, r) => (12, alloc::vec![*a, *r]), _ => return,
        };
        let mut addr: Vec<i64> = Vec::new();                                                  // remaining address list: (1, index) | (2, node id) pairs
        loop {
            match s.direction() { Address::Left => { addr.push(1); addr.push(0); } Address::Right => { addr.push(1); addr.push(1); }
                                  Address::Index(i) => { addr.push(1); addr.push(i as i64); } Address::Node(id) => { addr.push(2); addr.push(id.value() as i64); } Address::Null => break }
            s = s.peel();
        }
        unsafe { fdsp_bank_set(self.h, voice, kind, vals.as_ptr(), vals.len() as c_int, 0, addr.as_ptr(), (addr.len() / 2) as c_int); }
    }
    fn process(&mut self, size: usize, input: &BufferRef, output: &mut BufferMut) {           // audiounit.rs:45
        // BufferRef / BufferMut are [channel][64] f32, 32-byte aligned (buffer.rs:12,156) == the ABI layout
        let ip = if self.inputs > 0 { input.channel_f32(0).as_ptr() } else { core::ptr::null() };
        let rc = unsafe { fdsp_bank_process(self.h, size as u32, ip, output.channel_f32_mut(0).as_mut_ptr()) };
        if rc != 0 { self.failed = true; for c in 0..self.outputs { output.channel_f32_mut(c)[..size].fill(0.0); } }   // process() has no error channel
    }
    fn tick(&mut self, input: &[f32], output: &mut [f32]) {
        let mut i = crate::buffer::BufferVec::new(self.inputs.max(1)); let mut o = crate::buffer::BufferVec::new(self.outputs);
        for (c, x) in input.iter().enumerate() { i.set_f32(c, 0, *x); }
        self.process(1, &i.buffer_ref(), &mut o.buffer_mut());
        for (c, y) in output.iter_mut().enumerate() { *y = o.at_f32(c, 0); }
    }
    fn get_id(&self) -> u64 { 1000 }
    fn ping(&mut self, _probe: bool, hash: AttoHash) -> AttoHash { hash.hash(self.get_id()) }
    fn route(&mut self, input: &SignalFrame, _frequency: f64) -> SignalFrame { Routing::Arbitrary(0.0).route(input, self.outputs()) }
    fn footprint(&self) -> usize { core::mem::size_of::<Self>() }
}
impl Clone for GpuBank {                                                                      // dyn_clone (audiounit.rs:373): deep copy incl. device state
    fn clone(&self) -> Self {
        let mut b: *mut FdspBank = core::ptr::null_mut();
        let rc = unsafe { fdsp_bank_clone(self.h, &mut b) };
        GpuBank { h: b, inputs: self.inputs, outputs: self.outputs, failed: self.failed || rc != 0 }
    }
}
impl Drop for GpuBank { fn drop(&mut self) { if !self.h.is_null() { unsafe { fdsp_bank_destroy(self.h) } } } }

/// The ranks whose banks are mixed down together (one process per GPU): voices shard, the ONE exchange step is the sum of the per-GPU
/// mixes, below the C ABI (NCCL gather over NVLink + rank-order fold). `id` comes from rank 0 (`GpuGroup::unique_id`) and reaches the
/// other ranks however the host likes (a file, a socket, MPI).
pub struct GpuGroup { h: *mut FdspGroup, pub rank: i32, pub nranks: i32 }
unsafe impl Send for GpuGroup {}
impl GpuGroup {
    pub fn unique_id() -> Result<[u8; 128], String> { let mut id = [0u8; 128]; check(unsafe { fdsp_group_unique_id(id.as_mut_ptr() as *mut c_void, 128) })?; Ok(id) }
    pub fn new(nranks: i32, rank: i32, id: &[u8; 128], device: i32) -> Result<Self, String> {
        let mut g: *mut FdspGroup = core::ptr::null_mut();
        check(unsafe { fdsp_group_create(nranks, rank, id.as_ptr() as *const c_void, device, &mut g) })?;
        Ok(GpuGroup { h: g, rank, nranks })
    }
    /// Every rank renders its shard; rank `root` receives the finished mix in `out` ([outputs][samples]).
    pub fn render_reduced(&self, bank: &mut GpuBank, samples: usize, input: Option<&[f32]>, out: Option<&mut [f32]>, root: i32) -> Result<(), String> {
        let op = out.map_or(core::ptr::null_mut(), |o| o.as_mut_ptr());
        check(unsafe { fdsp_bank_render_reduced(bank.h, self.h, samples as u64, input.map_or(core::ptr::null(), |x| x.as_ptr()), op, root) })
    }
}
impl Drop for GpuGroup { fn drop(&mut self) { if !self.h.is_null() { unsafe { fdsp_group_destroy(self.h) } } } }
