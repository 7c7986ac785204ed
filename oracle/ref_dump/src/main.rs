//! Writes golden vectors of the reference crate for the parity tests (tests/test_ref_vectors.py): per BASELINE configuration a few
//! voices rendered with `Wave::render` semantics (block 64, 48 kHz) as little-endian f32 `[channels][samples]`, plus `manifest.json`.
//! Voice parameters are the ones fundsp_b200/workloads.py draws: u_k(i) = rnd1(4 i + k) (src/math.rs:569), so no numbers are shared
//! through files. Also per-node vectors (sine, saw, svf, moog, noise, reverb tail) that pin libm / wide / microfft at the ulp level.
use fundsp::prelude32::*;
use std::{fs, io::Write, path::Path};

const SR: f64 = 48000.0;

fn u(i: u64, k: u64) -> f64 { rnd1(4 * i + k) as f64 }
fn lerpf(a: f64, b: f64, t: f64) -> f64 { a * (1.0 - t) + b * t }
fn xerpf(a: f64, b: f64, t: f64) -> f64 { (lerpf(a.ln(), b.ln(), t)).exp() }

/// `AudioUnit::process` in 64-sample blocks like Wave::render (wave.rs:441-466); `gate`: optional single input channel.
fn render(unit: &mut dyn AudioUnit, n: usize, gate: Option<&[f32]>) -> Vec<Vec<f32>> {
    unit.set_sample_rate(SR);
    unit.allocate();
    let (ni, no) = (unit.inputs(), unit.outputs());
    let mut out = vec![vec![0.0f32; n]; no];
    let mut ib = BufferVec::new(ni.max(1));
    let mut ob = BufferVec::new(no);
    let mut t = 0;
    while t < n {
        let m = (n - t).min(64);
        if let Some(g) = gate { for i in 0..m { ib.set_f32(0, i, g[t + i]); } }
        unit.process(m, &ib.buffer_ref(), &mut ob.buffer_mut());
        for c in 0..no { for i in 0..m { out[c][t + i] = ob.at_f32(c, i); } }
        t += m;
    }
    out
}

fn dump(dir: &Path, name: &str, rows: &[Vec<f32>], manifest: &mut Vec<String>) {
    let mut f = fs::File::create(dir.join(format!("{name}.f32"))).unwrap();
    for r in rows { for x in r { f.write_all(&x.to_le_bytes()).unwrap(); } }
    manifest.push(format!("  {{\"name\": \"{}\", \"channels\": {}, \"samples\": {}}}", name, rows.len(), rows[0].len()));
}

fn gate(n: usize) -> Vec<f32> { (0..n).map(|t| if t >= 480 && t < 24000 { 1.0 } else { 0.0 }).collect() }

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "../../tests/golden/ref".into());
    let dir = Path::new(&dir);
    fs::create_dir_all(dir).unwrap();
    let mut m: Vec<String> = Vec::new();
    let n = 4800 + 61;   // 75 full blocks + one with a 5-sample tick-path tail
    // config 1
    dump(dir, "plumbing", &render(&mut (sine_hz(440.0) >> lowpass_hz(1000.0, 1.0)), 48000, None), &mut m);
    for i in [0u64, 5, 997, 16383] {
        // config 2: FM
        let f = xerpf(55.0, 1760.0, u(i, 0)) as f32; let md = lerpf(0.5, 8.0, u(i, 1)) as f32;
        dump(dir, &format!("fm_{i}"), &render(&mut (sine_hz(f).phase(u(i, 2) as f32) * f * md + f >> sine().phase(u(i, 3) as f32)), n, None), &mut m);
        // config 3a
        let fc = xerpf(100.0, 12000.0, u(i, 0)) as f32; let q = lerpf(0.5, 10.0, u(i, 1)) as f32;
        dump(dir, &format!("noise_svf_{i}"), &render(&mut (white().seed(i) >> lowpass_hz(fc, q)), n, None), &mut m);
        // headline
        let fs_ = xerpf(55.0, 1760.0, u(i, 2)) as f32;
        dump(dir, &format!("saw_svf_{i}"), &render(&mut (saw_hz(fs_).phase(u(i, 3) as f32) >> lowpass_hz(fc, q)), n, None), &mut m);
    }
    for i in [0u64, 5, 1023] {
        // config 4: subtractive voice with per-voice reverb; one input (gate), two outputs
        let f = xerpf(55.0, 880.0, u(i, 0)) as f32; let fc = xerpf(200.0, 8000.0, u(i, 1)) as f32; let q = lerpf(0.1, 0.9, u(i, 2)) as f32; let p = lerpf(-1.0, 1.0, u(i, 3)) as f32;
        let mut v = (((dc(f) >> saw()) | dc((fc, q))) >> moog()) * adsr_live(0.01, 0.1, 0.6, 0.3) >> pan(p) >> (multipass::<U2>() & 0.2 * reverb_stereo(10.0, 2.0, 0.5));
        let nn = 28800 + 7;
        dump(dir, &format!("subtractive_{i}"), &render(&mut v, nn, Some(&gate(nn))), &mut m);
    }
    // per-node pins at the libm / wide / microfft level
    dump(dir, "node_sine_440", &render(&mut sine_hz(440.0).phase(0.25), n, None), &mut m);
    dump(dir, "node_saw_110", &render(&mut saw_hz(110.0).phase(0.0), n, None), &mut m);
    dump(dir, "node_saw_7040", &render(&mut saw_hz(7040.0).phase(0.0), n, None), &mut m);
    dump(dir, "node_moog", &render(&mut (white().seed(1) >> moog_hz(1000.0, 0.7)), n, None), &mut m);
    dump(dir, "node_reverb", &render(&mut ((white().seed(2) | white().seed(3)) >> reverb_stereo(10.0, 2.0, 0.5)), 9600, None), &mut m);
    fs::write(dir.join("manifest.json"), format!("{{\"sample_rate\": {SR}, \"generator\": \"oracle/ref_dump (fundsp {})\", \"vectors\": [\n{}\n]}}\n", "0.23.0", m.join(",\n"))).unwrap();
    println!("wrote {} vectors to {}", m.len(), dir.display());
}
