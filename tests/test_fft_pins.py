"""FFTW3 is the one piece of arithmetic on the path that lives outside the reference's sources (not vendored,
CMakeLists.txt:63-66) and exists nowhere in this image.  Its r2c / c2r DEFINITION -- unnormalised, forward sign
-1, c2r = complex transform over the slow dimension first and a half-complex-to-real transform over the fast
dimension last (imaginary parts of the self-conjugate entries ignored) -- is restated by oracle.r2c / c2r with
pocketfft.  These tests pin that restatement to three independent things:
  1. numbers FFTW ITSELF produced: the double-precision REDFT00 / RODFT00 reference outputs SciPy ships for its own
     tests (scipy/fftpack/tests/fftw_double_ref.npz); a DCT-I / DST-I is the r2c of the even / odd extension, so they
     check magnitude, normalisation and the SIGN of the forward transform;
  2. Intel MKL's DFTI (torch.fft on the CPU), an independent FFT library, at the benchmark's own lengths, including
     c2r of spectra that are NOT Hermitian in their kx = 0 / Nyquist columns (what press_stag_array and convec feed it);
  3. (-m gpu) cuFFT, against the library's own kernels through the C ABI.
"""
import os

import numpy as np
import pytest

from helpers import O, rel


def _fftw_ref():
    import scipy.fftpack
    f = os.path.join(os.path.dirname(scipy.fftpack.__file__), "tests", "fftw_double_ref.npz")
    if not os.path.exists(f):
        pytest.skip("SciPy's FFTW reference data is not installed")
    return np.load(f)


def test_r2c_against_numbers_fftw_produced():
    d = _fftw_ref()
    checked = 0
    for n in d["sizes"]:
        n = int(n)
        x = np.linspace(0, n - 1, n)
        if n >= 2:
            # REDFT00 (DCT-I): Y_k = x_0 + (-1)^k x_{n-1} + 2 sum x_j cos(pi j k / (n-1)) = DFT of the even extension
            ext = np.concatenate([x, x[-2:0:-1]])
            m = ext.size
            row = np.zeros((1, 1, m + 2)); row[0, 0, :m] = ext
            spec = O.r2c(row, m).view(np.complex128)[0, 0]
            ref = d[f"dct_1_{n}"]
            assert np.abs(spec.real[:n] - ref).max() <= 1e-13 * np.abs(ref).max(), n
            assert np.abs(spec.imag).max() <= 1e-12 * np.abs(ref).max(), n
            checked += 1
        # RODFT00 (DST-I): Y_k = 2 sum x_j sin(pi (j+1)(k+1)/(n+1)); DFT of the odd extension = -i Y (forward sign -1)
        ext = np.concatenate([[0.0], x, [0.0], -x[::-1]])
        m = ext.size
        row = np.zeros((1, 1, m + 2)); row[0, 0, :m] = ext
        spec = O.r2c(row, m).view(np.complex128)[0, 0]
        ref = d[f"dst_1_{n}"]
        assert np.abs(-spec.imag[1:n + 1] - ref).max() <= 1e-13 * np.abs(ref).max(), n
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("nx,ny", [(16, 16), (48, 32), (80, 96), (128, 192), (512, 512), (768, 768), (1024, 512), (1536, 768)])
def test_r2c_c2r_against_mkl(nx, ny):
    import torch
    if not torch.backends.mkl.is_available():
        pytest.skip("torch without MKL")
    rng = np.random.default_rng(nx + ny)
    ld = nx + 2
    a = np.zeros((2, ny, ld)); a[:, :, :nx] = rng.standard_normal((2, ny, nx))
    mine = O.r2c(a, nx).view(np.complex128)
    mkl = torch.fft.rfft2(torch.from_numpy(a[:, :, :nx].copy()), dim=(-2, -1)).numpy()
    assert rel(mine, mkl) <= 1e-15 * np.log2(nx * ny) * 4
    # c2r of a spectrum with arbitrary (non-Hermitian) kx = 0 and kx = nx/2 columns
    spec = rng.standard_normal((2, ny, ld // 2)) + 1j * rng.standard_normal((2, ny, ld // 2))
    back = O.c2r(spec.view(np.float64).copy(), nx)[:, :, :nx]
    t = torch.from_numpy(spec.copy())
    ref = torch.fft.irfft(torch.fft.ifft(t, dim=-2, norm="forward"), n=nx, dim=-1, norm="forward").numpy()
    assert rel(back, ref) <= 1e-15 * np.log2(nx * ny) * 4
    # ... which is: imaginary parts of the two self-conjugate columns dropped AFTER the y transform
    ycol = np.fft.ifft(spec, axis=1) * ny
    ycol[:, :, 0] = ycol[:, :, 0].real; ycol[:, :, -1] = ycol[:, :, -1].real
    assert rel(back, np.fft.irfft(ycol, n=nx, axis=2) * nx) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny", [(64, 48), (512, 512), (1024, 512)])
def test_cuda_fft_against_cufft(nx, ny):
    """The library's r2c / c2r (both grids) against cuFFT -- an independent implementation, used only as a checker."""
    import torch
    import lesgo_b200
    from helpers import make_dims
    p = O.Params(nx=nx, ny=ny, Nz=2)
    c = lesgo_b200.Core(make_dims(p, device=0))
    rng = np.random.default_rng(5)
    for big, n0, n1 in ((False, p.nx, p.ny), (True, p.nx2, p.ny2)):
        a = np.zeros((2, n1, n0 + 2)); a[:, :, :n0] = rng.standard_normal((2, n1, n0))
        got = c.fft_r2c(a.copy(), big=big).view(np.complex128)
        ref = torch.fft.rfft2(torch.from_numpy(a[:, :, :n0].copy()).cuda(), dim=(-2, -1)).cpu().numpy()
        assert rel(got, ref) <= 2e-14
        spec = rng.standard_normal((2, n1, n0 // 2 + 1)) + 1j * rng.standard_normal((2, n1, n0 // 2 + 1))
        back = c.fft_c2r(spec.view(np.float64).copy(), big=big)[:, :, :n0]
        t = torch.from_numpy(spec.copy()).cuda()
        y = torch.fft.ifft(t, dim=-2, norm="forward")
        y[:, :, 0] = y[:, :, 0].real.to(y.dtype); y[:, :, -1] = y[:, :, -1].real.to(y.dtype)
        ref = torch.fft.irfft(y, n=n0, dim=-1, norm="forward").cpu().numpy()
        assert rel(back, ref) <= 2e-14
