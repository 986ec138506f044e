//! Writes the reference's own outputs as a fixture: `cargo run --release --example dump_golden -- OUT.bin`.
//! `pytest tests/test_oracle.py` checks the C oracle against the file when it exists (tests/golden/rust_reference.bin),
//! which pins the oracle to the Rust implementation without needing Rust at test time.  No GPU involved.
//!
//! Format (little-endian): "GGQR", u32 version = 1, then records
//!   u32 kind (0 quantize, 1 dequantize), u32 ggml type, u32 float dtype (0 f32, 1 f16, 30 bf16), u64 n_elems,
//!   u64 in_bytes, u64 out_bytes, input bytes, output bytes.
use ggml_quants::{bf16, f16, DataBlock, QuantExt, Quantize};
use ggml_quants::{Q4_0, Q4_1, Q5_0, Q5_1, Q8K, Q8_0, Q8_1};
use std::io::Write;

struct Rng(u64);
impl Rng {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
    fn unit(&mut self) -> f64 { ((self.next() >> 11) as f64 + 0.5) / (1u64 << 53) as f64 }
    fn gauss(&mut self) -> f32 {
        let (u, v) = (self.unit(), self.unit());
        ((-2.0 * u.ln()).sqrt() * (2.0 * std::f64::consts::PI * v).cos()) as f32
    }
}

fn bytes<T>(v: &[T]) -> &[u8] { unsafe { std::slice::from_raw_parts(v.as_ptr().cast(), std::mem::size_of_val(v)) } }

fn inputs(n: usize, seed: u64) -> Vec<f32> {
    let mut rng = Rng(seed);
    let mut x: Vec<f32> = (0..n * 300).map(|_| rng.gauss() * 0.02).collect();
    x.extend((0..n * 20).map(|_| { let d = rng.gauss(); rng.gauss() / if d.abs() < 1e-3 { 1e-3 } else { d } }));
    // edge rows: zeros, -zeros, constant, +-tie, NaN, inf, denormal, huge
    let z = vec![0.0f32; n];
    x.extend(&z);
    x.extend(z.iter().map(|v| -v));
    x.extend(vec![0.37f32; n]);
    let mut a = z.clone(); a[5 % n] = 1.0; a[20 % n] = -1.0; x.extend(&a);
    let mut a = z.clone(); a[5 % n] = -1.0; a[20 % n] = 1.0; x.extend(&a);
    let mut a: Vec<f32> = (0..n).map(|_| rng.gauss()).collect(); a[3 % n] = f32::NAN; x.extend(&a);
    let mut a: Vec<f32> = (0..n).map(|_| rng.gauss()).collect(); a[7 % n] = f32::INFINITY; x.extend(&a);
    x.extend((0..n).map(|_| rng.gauss() * 1e-41));
    x.extend((0..n).map(|_| rng.gauss() * 1e6));
    let mut a = z.clone(); a[1 % n] = -0.0; a[2 % n] = 1.0; x.extend(&a);
    x
}

fn record(out: &mut impl Write, kind: u32, ty: u32, fdt: u32, n: usize, input: &[u8], output: &[u8]) {
    out.write_all(&kind.to_le_bytes()).unwrap();
    out.write_all(&ty.to_le_bytes()).unwrap();
    out.write_all(&fdt.to_le_bytes()).unwrap();
    out.write_all(&(n as u64).to_le_bytes()).unwrap();
    out.write_all(&(input.len() as u64).to_le_bytes()).unwrap();
    out.write_all(&(output.len() as u64).to_le_bytes()).unwrap();
    out.write_all(input).unwrap();
    out.write_all(output).unwrap();
}

fn dump<Blk, T, const N: usize>(out: &mut impl Write, ty: u32, fdt: u32, conv: impl Fn(f32) -> T, zero: T)
where
    Blk: Quantize<T, N> + Send + Sync,
    T: Copy + Send + Sync,
{
    let x: Vec<T> = inputs(N, 1000 + ty as u64).into_iter().map(conv).collect();
    let mut q: Vec<Blk> = (0..x.len() / N).map(|_| Blk::ZEROS).collect();
    <Blk as QuantExt<T, N>>::quantize_slice(&mut q, &x).unwrap();
    record(out, 0, ty, fdt, x.len(), bytes(&x), bytes(&q));
    // dequantize: the quantized blocks plus random bytes (wild scales, every code value)
    let mut rng = Rng(2000 + ty as u64);
    let mut wild: Vec<Blk> = (0..257).map(|_| Blk::ZEROS).collect();
    unsafe { std::slice::from_raw_parts_mut(wild.as_mut_ptr().cast::<u8>(), 257 * std::mem::size_of::<Blk>()) }.iter_mut().for_each(|b| *b = rng.next() as u8);
    for blocks in [&q, &wild] {
        let mut y = vec![zero; blocks.len() * N];
        <Blk as QuantExt<T, N>>::dequantize_slice(&mut y, blocks).unwrap();
        record(out, 1, ty, fdt, y.len(), bytes(blocks), bytes(&y));
    }
}

fn main() {
    let path = std::env::args().nth(1).expect("usage: dump_golden OUT.bin");
    let mut out = std::io::BufWriter::new(std::fs::File::create(path).unwrap());
    out.write_all(b"GGQR").unwrap();
    out.write_all(&1u32.to_le_bytes()).unwrap();
    macro_rules! all { ($blk:ty, $ty:expr, $n:expr) => {
        dump::<$blk, f32, $n>(&mut out, $ty, 0, |v| v, 0.0);
        dump::<$blk, f16, $n>(&mut out, $ty, 1, f16::from_f32, f16::ZERO);
        dump::<$blk, bf16, $n>(&mut out, $ty, 30, bf16::from_f32, bf16::ZERO);
    }; }
    all!(Q4_0, 2, 32);
    all!(Q4_1, 3, 32);
    all!(Q5_0, 6, 32);
    all!(Q5_1, 7, 32);
    all!(Q8_0, 8, 32);
    all!(Q8_1, 9, 32);
    all!(Q8K, 15, 256);
    dump::<f16, f32, 1>(&mut out, 1, 0, |v| v, 0.0);
    dump::<bf16, f32, 1>(&mut out, 30, 0, |v| v, 0.0);
}
