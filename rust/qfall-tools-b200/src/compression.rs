//! `impl LossyCompressionFIPS203` (lossy_compression_fips203.rs:20-59) backed by the GPU kernels.
//!
//! The trait and the qfall-math types both live in other crates, so the orphan rule forbids
//! `impl LossyCompressionFIPS203 for PolynomialRingZq` here; the shim provides transparent newtypes instead
//! ([`PolynomialRingZqB200`], [`MatPolynomialRingZqB200`]) with the reference's associated types
//! (`CompressedType = PolyOverZ / MatPolyOverZ`, `ModulusType = ModulusPolynomialRingZq`, :61-63, :175-177).
//! Inside qfall-tools itself the two `impl` blocks can call [`crate::compress_words`] directly.
//! Coefficients cross the boundary as FLINT small words (`i64`, q < 2^62): one `qf_compress_i64` call per value,
//! however many polynomials it holds.  Shipped as source (see lib.rs).
use crate::{compress_words, decompress_words};
use qfall_math::{
    integer::{MatPolyOverZ, PolyOverZ, Z},
    integer_mod_q::{MatPolynomialRingZq, ModulusPolynomialRingZq, PolynomialRingZq},
    traits::{GetCoefficient, MatrixDimensions, MatrixGetEntry, MatrixSetEntry, SetCoefficient},
};
use qfall_tools::compression::LossyCompressionFIPS203;

#[repr(transparent)]
pub struct PolynomialRingZqB200(pub PolynomialRingZq);
#[repr(transparent)]
pub struct MatPolynomialRingZqB200(pub MatPolynomialRingZq);

fn d_to_u32(d: impl Into<Z>) -> u32 {
    let d: Z = d.into();
    // lossy_compression_fips203.rs:91-94 / :149-152
    assert!(d >= Z::ONE, "Performing this function with d < 1 implies reducing mod 1");
    u32::try_from(&d).expect("d below 2^32")
}

fn q_of(modulus: &ModulusPolynomialRingZq) -> u64 {
    u64::try_from(&modulus.get_q()).expect("modulus below 2^62")
}

/// coefficients 0..=deg of the least non-negative representative (the reference loops to the actual degree, :101)
fn poly_words(p: &PolyOverZ) -> Vec<i64> {
    (0..=p.get_degree()).map(|i| i64::try_from(&GetCoefficient::<Z>::get_coeff(p, i).unwrap()).unwrap()).collect()
}

fn poly_from_words(w: &[i64]) -> PolyOverZ {
    let mut out = PolyOverZ::default();
    for (i, v) in w.iter().enumerate() {
        if *v != 0 {
            out.set_coeff(i as i64, Z::from(*v)).unwrap();
        }
    }
    out
}

impl LossyCompressionFIPS203 for PolynomialRingZqB200 {
    type CompressedType = PolyOverZ;
    type ModulusType = ModulusPolynomialRingZq;

    /// lossy_compression_fips203.rs:89-114
    fn lossy_compress(&self, d: impl Into<Z>) -> PolyOverZ {
        let d = d_to_u32(d);
        let q = q_of(&self.0.get_mod());
        let words = poly_words(&self.0.get_representative_least_nonnegative_residue());
        poly_from_words(&compress_words(&words, d, q))
    }

    /// lossy_compression_fips203.rs:143-172
    fn lossy_decompress(compressed: &PolyOverZ, d: impl Into<Z>, modulus: &ModulusPolynomialRingZq) -> Self {
        let d = d_to_u32(d);
        let words = decompress_words(&poly_words(compressed), d, q_of(modulus));
        PolynomialRingZqB200(PolynomialRingZq::from((&poly_from_words(&words), modulus)))
    }
}

impl LossyCompressionFIPS203 for MatPolynomialRingZqB200 {
    type CompressedType = MatPolyOverZ;
    type ModulusType = ModulusPolynomialRingZq;

    /// lossy_compression_fips203.rs:203-217: entrywise, here ONE device call for the whole matrix
    fn lossy_compress(&self, d: impl Into<Z>) -> MatPolyOverZ {
        let d = d_to_u32(d);
        let q = q_of(&self.0.get_mod());
        let lifted = self.0.get_representative_least_nonnegative_residue();
        let (rows, cols) = (lifted.get_num_rows(), lifted.get_num_columns());
        let mut words = Vec::new();
        let mut lens = Vec::with_capacity((rows * cols) as usize);
        for i in 0..rows {
            for j in 0..cols {
                let p: PolyOverZ = lifted.get_entry(i, j).unwrap();
                let w = poly_words(&p);
                lens.push(w.len());
                words.extend(w);
            }
        }
        let comp = compress_words(&words, d, q);
        let mut out = MatPolyOverZ::new(rows, cols);
        let mut off = 0;
        for i in 0..rows {
            for j in 0..cols {
                let len = lens[(i * cols + j) as usize];
                out.set_entry(i, j, &poly_from_words(&comp[off..off + len])).unwrap();
                off += len;
            }
        }
        out
    }

    /// lossy_compression_fips203.rs:246-268
    fn lossy_decompress(compressed: &MatPolyOverZ, d: impl Into<Z>, modulus: &ModulusPolynomialRingZq) -> Self {
        let d = d_to_u32(d);
        let q = q_of(modulus);
        let (rows, cols) = (compressed.get_num_rows(), compressed.get_num_columns());
        let mut words = Vec::new();
        let mut lens = Vec::with_capacity((rows * cols) as usize);
        for i in 0..rows {
            for j in 0..cols {
                let p: PolyOverZ = compressed.get_entry(i, j).unwrap();
                let w = poly_words(&p);
                lens.push(w.len());
                words.extend(w);
            }
        }
        let dec = decompress_words(&words, d, q);
        let mut out = MatPolyOverZ::new(rows, cols);
        let mut off = 0;
        for i in 0..rows {
            for j in 0..cols {
                let len = lens[(i * cols + j) as usize];
                out.set_entry(i, j, &poly_from_words(&dec[off..off + len])).unwrap();
                off += len;
            }
        }
        MatPolynomialRingZqB200(MatPolynomialRingZq::from((&out, modulus)))
    }
}
