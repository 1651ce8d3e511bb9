"""common_encodings.rs (second half): the oracle against the reference's own unit tests (CPU), and the
device kernels (through the C ABI / host mirror) bit-exact against the oracle (GPU)."""
import numpy as np
import pytest

from oracle import qfall_oracle as O


def test_oracle_encode_reference_cases():
    # common_encodings.rs:163-186 (binary), :189-214 (ternary): q = 257, X^16 + 1
    q = 257
    assert O.encode_value_in_polynomialringzq(1, 2, 16, q) == [128] + [0] * 15
    assert O.encode_value_in_polynomialringzq(2, 2, 16, q) == [0, 128] + [0] * 14
    assert O.encode_value_in_polynomialringzq(3, 2, 16, q) == [128, 128] + [0] * 14
    assert O.encode_value_in_polynomialringzq(1, 3, 16, q) == [85] + [0] * 15
    assert O.encode_value_in_polynomialringzq(2, 3, 16, q) == [170] + [0] * 15
    assert O.encode_value_in_polynomialringzq(3, 3, 16, q) == [0, 85] + [0] * 14
    # :217-246: not enough space, base < 2, negative value
    for args in ((65536, 2, 16, q), (4, 1, 16, q), (-1, 1, 16, q), (-1, 2, 16, q)):
        with pytest.raises(ValueError):
            O.encode_value_in_polynomialringzq(*args)
    assert O.encode_value_in_polynomialringzq(65535, 2, 16, q) == [128] * 16
    assert O.encode_value_in_polynomialringzq(0, 2, 16, q) == [0] * 16


def test_oracle_round_trip_reference_cases():
    # common_encodings.rs:258-285 (X^17 + 1 binary, X^16 + 1 ternary), :288-296
    rng = np.random.default_rng(0)
    for base, n in ((2, 17), (3, 16)):
        for msg in [0, 1, 65535] + [int(x) for x in rng.integers(0, 65535, 50)]:
            enc = O.encode_value_in_polynomialringzq(msg, base, n, 257)
            assert O.decode_value_from_polynomialringzq(enc, base, 257) == msg
    with pytest.raises(ValueError):
        O.decode_value_from_polynomialringzq([1] * 16, 1, 257)
    # the formula adds floor(q / (2 base)) to c * base before dividing by q (:131,145-146): digit d is recovered from
    # d q / base + x for -q / (2 base^2) <= x < q / base - q / (2 base^2)
    enc = O.encode_value_in_polynomialringzq(0b1011, 2, 16, 3329)
    noisy = [(c + e) % 3329 for c, e in zip(enc, [1200, -415, 1247, -400] + [-415] * 12)]
    assert O.decode_value_from_polynomialringzq(noisy, 2, 3329) == 0b1011
    assert O.decode_value_from_polynomialringzq([(enc[0] - 416) % 3329] + enc[1:], 2, 3329) != 0b1011


@pytest.mark.gpu
@pytest.mark.parametrize("q,base,n", [(257, 2, 16), (257, 3, 16), (3329, 2, 256), (3329, 5, 64), (2**31 - 1, 7, 33),
                                      (2**61 - 1, 256, 20), (65521, 255, 16)])
def test_device_encodings_match_oracle(q, base, n):
    from tools_b200 import encodings as E

    rng = np.random.default_rng(n)
    top = base**n
    vals = [0, 1, top - 1] + [int(rng.integers(0, 2**62)) % top for _ in range(40)]
    enc = E.encode_values_batch(vals, base, n, q)
    assert enc.shape == (len(vals), n)
    for v, row in zip(vals, enc):
        assert row.astype(object).tolist() == O.encode_value_in_polynomialringzq(v, base, n, q)
    # decode(encode(v)) = v needs (base - 1) (q mod base) <= floor(q / (2 base)) -- the spread is floor(q / base), not
    # q / base; otherwise the reference's own formulas do not round-trip (parity with the oracle still holds)
    round_trips = (base - 1) * (q % base) <= q // (2 * base)
    dec = E.decode_values_batch(enc, base, q)
    assert dec == [O.decode_value_from_polynomialringzq(r.astype(object).tolist(), base, q) for r in enc]
    assert (dec == vals) == round_trips or not round_trips
    # noisy coefficients, any representative (negative, above q): decode digit by digit against the oracle
    tol = max(0, q // (2 * base * base) - 1)  # see test_oracle_round_trip_reference_cases
    for amp, recover in ((tol // 2, True), (2 * tol + 3, False)):  # inside the decoding radius / beyond it (parity only)
        noise = rng.integers(-amp, amp + 1, enc.shape)
        shifted = enc.astype(object) + noise.astype(object) + rng.integers(-2, 3, enc.shape).astype(object) * q
        if q < 2**40:
            got = E.decode_values_batch(np.asarray(shifted, dtype=np.int64), base, q)
            assert got == [O.decode_value_from_polynomialringzq(r.tolist(), base, q) for r in shifted]
            if recover and round_trips and tol > 1:
                assert got == vals
    # trait-shaped single calls and the reference's error cases
    one = E.encode_value_in_polynomialringzq(vals[5], base, n, q)
    assert E.decode_value_from_polynomialringzq(one, base, q) == O.decode_value_from_polynomialringzq(one.astype(object).tolist(), base, q)
    with pytest.raises(E.MathError):
        E.encode_value_in_polynomialringzq(top, base, n, q)
    with pytest.raises(E.MathError):
        E.encode_value_in_polynomialringzq(-1, base, n, q)
    with pytest.raises(E.MathError):
        E.encode_value_in_polynomialringzq(1, 1, n, q)
    with pytest.raises(E.MathError):
        E.decode_value_from_polynomialringzq(one, 1, q)


@pytest.mark.gpu
def test_device_bit_packed_messages():
    """32-byte messages <-> 256 coefficients (base 2), ragged counts, against the oracle's per-value functions."""
    from tools_b200 import encodings as E

    q = 3329
    rng = np.random.default_rng(5)
    for count in (1, 3, 1000):
        msg = rng.integers(0, 256, (count, 32), dtype=np.uint8)
        c = E.encode_bits(msg, q)
        assert c.shape == (count, 256) and c.dtype == np.uint16
        for i in range(min(count, 5)):
            value = int.from_bytes(bytes(msg[i]), "little")
            assert c[i].tolist() == O.encode_value_in_polynomialringzq(value, 2, 256, q)
        noisy = ((c.astype(np.int64) + rng.integers(-415, 416, c.shape)) % q).astype(np.uint16)
        back = E.decode_bits(noisy, q)
        assert np.array_equal(back, msg)
        assert int.from_bytes(bytes(back[0]), "little") == O.decode_value_from_polynomialringzq(noisy[0].tolist(), 2, q)
