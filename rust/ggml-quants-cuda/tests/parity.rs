//! Byte parity of libggq against the UNMODIFIED reference crate on identical inputs.
//!
//! For every block type whose arithmetic the reference implements (Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1, Q8K and the
//! f16 / bf16 "blocks") and every float side T in {f32, f16, bf16}:
//!   * `QuantExt::quantize_slice` (CPU, ggml-quants/src/lib.rs:121-133) and `CudaQuantExt::quantize_slice` (GPU)
//!     must write the same bytes for Gaussian, heavy-tailed and edge-case blocks (zeros, ties, NaN, inf, denormals);
//!   * `dequantize_slice` of random packed bytes must give the same bits (a NaN produced by arithmetic on a
//!     non-finite scale may differ in payload between x86 and the GPU; both sides NaN counts as equal and the
//!     number of such elements is printed).
//! The K-quants (Q2K..Q6K) are `todo!()` in the reference, so only their GPU round trip is bounded here.
//! Needs a CUDA device: libggq has no CPU fallback.
use ggml_quants::{bf16, f16, DataBlock, QuantExt, Quantize, QuantizeError};
use ggml_quants::{Q2K, Q3K, Q4K, Q4_0, Q4_1, Q5K, Q5_0, Q5_1, Q6K, Q8K, Q8_0, Q8_1};
use ggml_quants_cuda::{bytes_of, CudaQuantExt, FloatSide, GgmlBlock};

/// SplitMix64: the whole test is reproducible from the seeds below.
struct Rng(u64);
impl Rng {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
    fn unit(&mut self) -> f64 {
        ((self.next() >> 11) as f64 + 0.5) / (1u64 << 53) as f64
    }
    fn gauss(&mut self) -> f32 {
        let (u, v) = (self.unit(), self.unit());
        ((-2.0 * u.ln()).sqrt() * (2.0 * std::f64::consts::PI * v).cos()) as f32
    }
}

trait Side: FloatSide + PartialEq + std::fmt::Debug {
    fn from32(v: f32) -> Self;
    fn to32(self) -> f32;
    fn zero() -> Self;
}
impl Side for f32 {
    fn from32(v: f32) -> Self { v }
    fn to32(self) -> f32 { self }
    fn zero() -> Self { 0.0 }
}
impl Side for f16 {
    fn from32(v: f32) -> Self { f16::from_f32(v) }
    fn to32(self) -> f32 { self.to_f32() }
    fn zero() -> Self { f16::ZERO }
}
impl Side for bf16 {
    fn from32(v: f32) -> Self { bf16::from_f32(v) }
    fn to32(self) -> f32 { self.to_f32() }
    fn zero() -> Self { bf16::ZERO }
}

/// The rows of tests/data.py::edge_blocks, for a block of `n` elements.
fn edge_blocks(n: usize, rng: &mut Rng) -> Vec<f32> {
    let z = vec![0.0f32; n];
    let mut rows: Vec<Vec<f32>> = vec![z.clone(), z.iter().map(|v| -v).collect(), vec![0.37; n], vec![-2.5; n]];
    let put = |pairs: &[(usize, f32)]| { let mut a = z.clone(); for &(i, v) in pairs { a[i % n] = v; } a };
    rows.push(put(&[(5, 1.0), (20, -1.0)])); // |x| tie: the first wins (+)
    rows.push(put(&[(5, -1.0), (20, 1.0)])); // (-)
    rows.push(put(&[(n - 1, 3.0)]));
    rows.push(put(&[(0, -3.0)]));
    let mut g = |scale: f32| -> Vec<f32> { (0..n).map(|_| rng.gauss() * scale).collect() };
    let mut a = g(1.0); a[3] = f32::NAN; rows.push(a);
    rows.push(vec![f32::NAN; n]);
    let mut a = g(1.0); a[7] = f32::INFINITY; rows.push(a);
    let mut a = g(1.0); a[9] = f32::NEG_INFINITY; rows.push(a);
    let mut a = g(1.0); a[1] = f32::INFINITY; a[2] = f32::NEG_INFINITY; rows.push(a);
    rows.push(g(1e-41)); // f32 denormals
    rows.push(put(&[(4, 1e-45)])); // delta underflows to 0
    rows.push(put(&[(4, -1e-45), (6, 1e-45)]));
    rows.push(g(6e-8)); // f16-subnormal scales
    rows.push(g(3e4)); // near f16 overflow
    rows.push(g(1e6)); // delta overflows f16 -> inf
    rows.push(put(&[(0, 0.0), (1, -0.0), (2, 1.0)])); // +0 first among the minima
    rows.push(put(&[(0, -0.0), (1, 0.0), (2, 1.0)])); // -0 first
    rows.push((0..n).map(|i| -1.0 + 2.0 * i as f32 / (n - 1) as f32).collect()); // exact .5 rounding cases
    rows.push((0..n).map(|i| (i as f32 - n as f32 / 2.0) * 0.5).collect());
    rows.push((1..=n).map(|i| i as f32 * 0.1).collect());
    rows.concat()
}

fn inputs(n: usize, seed: u64) -> Vec<f32> {
    let mut rng = Rng(seed);
    let mut x: Vec<f32> = (0..n * 1201).map(|_| rng.gauss() * 0.02).collect();
    // heavy tails: ratio of Gaussians
    x.extend((0..n * 37).map(|_| { let d = rng.gauss(); rng.gauss() / if d.abs() < 1e-3 { 1e-3 } else { d } * 0.02 }));
    x.extend((0..n * 13).map(|_| rng.unit() as f32));
    x.extend(edge_blocks(n, &mut rng));
    x
}

fn zeros<B: DataBlock>(n: usize) -> Vec<B> {
    (0..n).map(|_| B::ZEROS).collect()
}

/// f16 header fields per block type (byte offsets): a NaN there compares equal to any NaN.
fn f16_fields(ty: u32) -> &'static [usize] {
    match ty { 2 | 6 | 8 | 15 => &[0], 3 | 7 | 9 => &[0, 2], 10 => &[80, 82], 11 => &[108], 12 | 13 => &[0, 2], 14 => &[208], _ => &[] }
}

fn assert_same_blocks(ty: u32, size: usize, cpu: &[u8], gpu: &[u8], what: &str) {
    assert_eq!(cpu.len(), gpu.len());
    let mut nan_fields = 0usize;
    for (b, (c, g)) in cpu.chunks(size).zip(gpu.chunks(size)).enumerate() {
        if c == g { continue; }
        let (mut c, mut g) = (c.to_vec(), g.to_vec());
        for &o in f16_fields(ty) {
            let (fc, fg) = (u16::from_le_bytes([c[o], c[o + 1]]), u16::from_le_bytes([g[o], g[o + 1]]));
            if fc & 0x7FFF > 0x7C00 && fg & 0x7FFF > 0x7C00 { g[o] = c[o]; g[o + 1] = c[o + 1]; nan_fields += 1; }
        }
        assert_eq!(c, g, "{what}: block {b} differs");
    }
    if nan_fields > 0 { println!("{what}: {nan_fields} NaN header fields compared as NaN ~ NaN"); }
}

fn parity<Blk, T, const N: usize>(name: &str)
where
    Blk: Quantize<T, N> + GgmlBlock + Send + Sync,
    T: Side,
{
    let size = std::mem::size_of::<Blk>();
    // ---- quantize: same bytes ----
    let x: Vec<T> = inputs(N, 100 + Blk::TY as u64).into_iter().map(T::from32).collect();
    let nb = x.len() / N;
    let (mut cpu, mut gpu) = (zeros::<Blk>(nb), zeros::<Blk>(nb));
    <Blk as QuantExt<T, N>>::quantize_slice(&mut cpu, &x).unwrap();
    <Blk as CudaQuantExt<T>>::quantize_slice(&mut gpu, &x).unwrap();
    assert_same_blocks(Blk::TY, size, bytes_of(&cpu), bytes_of(&gpu), &format!("{name} quantize fdt {}", T::FDT));
    // ---- dequantize: same bits, on the quantized blocks and on random bytes (every code value, wild scales) ----
    let mut rng = Rng(200 + Blk::TY as u64);
    let mut wild = zeros::<Blk>(2051);
    unsafe { std::slice::from_raw_parts_mut(wild.as_mut_ptr().cast::<u8>(), 2051 * size) }.iter_mut().for_each(|b| *b = rng.next() as u8);
    for (label, blocks) in [("quantized", &cpu), ("random bytes", &wild)] {
        let n = blocks.len() * N;
        let (mut yc, mut yg) = (vec![T::zero(); n], vec![T::zero(); n]);
        <Blk as QuantExt<T, N>>::dequantize_slice(&mut yc, blocks).unwrap();
        <Blk as CudaQuantExt<T>>::dequantize_slice(&mut yg, blocks).unwrap();
        let (mut both_nan, mut first_bad) = (0usize, None);
        for (i, (a, b)) in yc.iter().zip(&yg).enumerate() {
            if bytes_of(std::slice::from_ref(a)) == bytes_of(std::slice::from_ref(b)) { continue; }
            if a.to32().is_nan() && b.to32().is_nan() { both_nan += 1; } else if first_bad.is_none() { first_bad = Some(i); }
        }
        assert!(first_bad.is_none(), "{name} dequantize fdt {} ({label}): element {:?} differs", T::FDT, first_bad);
        if both_nan > 0 { println!("{name} dequantize fdt {} ({label}): {both_nan} of {n} elements are NaN on both sides with different payloads", T::FDT); }
    }
    // ---- the two length checks, in the reference's order (lib.rs:293-331) ----
    let mut one = zeros::<Blk>(1);
    let mut three = zeros::<Blk>(3);
    let short = vec![T::zero(); N.max(2) - 1];
    let two = vec![T::zero(); 2 * N];
    if N > 1 {
        assert_eq!(<Blk as CudaQuantExt<T>>::quantize_slice(&mut one, &short), Err(QuantizeError::Indivisible));
        assert_eq!(<Blk as CudaQuantExt<T>>::quantize_slice(&mut three, &short), Err(QuantizeError::Indivisible));
    }
    assert_eq!(<Blk as CudaQuantExt<T>>::quantize_slice(&mut three, &two), Err(QuantizeError::LengthMismatch));
    let mut out = vec![T::zero(); 2 * N];
    assert_eq!(<Blk as CudaQuantExt<T>>::dequantize_slice(&mut out, &three), Err(QuantizeError::LengthMismatch));
}

macro_rules! parity_tests {
    ($($name:ident: $blk:ty, $n:expr;)*) => { $(
        #[test]
        fn $name() {
            parity::<$blk, f32, $n>(stringify!($blk));
            parity::<$blk, f16, $n>(stringify!($blk));
            parity::<$blk, bf16, $n>(stringify!($blk));
        }
    )* };
}
parity_tests! {
    q4_0: Q4_0, 32; q4_1: Q4_1, 32; q5_0: Q5_0, 32; q5_1: Q5_1, 32; q8_0: Q8_0, 32; q8_1: Q8_1, 32; q8_k: Q8K, 256;
}

/// f16 / bf16 as 1-element blocks (structs/half.rs:8-38): every f32 -> half conversion of the inputs, and all 65 536
/// half patterns widened, NaN payloads included (casts must be bit-exact, no NaN ~ NaN rule here).
#[test]
fn half_blocks() {
    let x = inputs(32, 7);
    macro_rules! one { ($h:ty) => {{
        let (mut c, mut g) = (vec![<$h>::ZERO; x.len()], vec![<$h>::ZERO; x.len()]);
        <$h as QuantExt<f32, 1>>::quantize_slice(&mut c, &x).unwrap();
        <$h as CudaQuantExt<f32>>::quantize_slice(&mut g, &x).unwrap();
        assert_eq!(bytes_of(&c), bytes_of(&g));
        let all: Vec<$h> = (0..=u16::MAX).map(<$h>::from_bits).collect();
        let (mut c, mut g) = (vec![0f32; all.len()], vec![0f32; all.len()]);
        <$h as QuantExt<f32, 1>>::dequantize_slice(&mut c, &all).unwrap();
        <$h as CudaQuantExt<f32>>::dequantize_slice(&mut g, &all).unwrap();
        assert_eq!(bytes_of(&c), bytes_of(&g));
    }}; }
    one!(f16);
    one!(bf16);
}

/// K-quants: `todo!()` in the reference (structs/q{2..6}_k.rs), upstream-ggml arithmetic in libggq.  Only a GPU round
/// trip can be checked from here; the bound is the rmse upstream's own test-quantize-fns accepts for each type.
#[test]
fn k_quants_round_trip() {
    fn rt<Blk: GgmlBlock + DataBlock>(max_rmse: f32) {
        let mut rng = Rng(300 + Blk::TY as u64);
        let x: Vec<f32> = (0..256 * 512).map(|_| rng.gauss()).collect();
        let mut q = zeros::<Blk>(512);
        <Blk as CudaQuantExt<f32>>::quantize_slice(&mut q, &x).unwrap();
        let mut y = vec![0f32; x.len()];
        <Blk as CudaQuantExt<f32>>::dequantize_slice(&mut y, &q).unwrap();
        let mse: f64 = x.iter().zip(&y).map(|(a, b)| ((a - b) as f64).powi(2)).sum::<f64>() / x.len() as f64;
        assert!((mse.sqrt() as f32) < max_rmse, "type {} rmse {}", Blk::TY, mse.sqrt());
    }
    rt::<Q2K>(0.35);
    rt::<Q3K>(0.2);
    rt::<Q4K>(0.1);
    rt::<Q5K>(0.05);
    rt::<Q6K>(0.03);
}
