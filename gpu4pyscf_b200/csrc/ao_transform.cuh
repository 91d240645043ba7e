// Sparse AO-basis transforms between the molecule's (contracted, spherical or Cartesian) AO basis
// and the sorted, decontracted Cartesian basis the J/K kernels work in:
//     D_sorted = C D C^T          V = C^T V_sorted C
// C (nao_sorted x nao) holds the recontraction coefficients times the Cartesian->spherical maps and
// has a handful of non-zeros per row.  Play the role of the reference's bra_/ket_ sph2sorted /
// sorted2sph / cart2sorted / sorted2cart kernels (lib/gvhf-rys/mole_helper.cu:1735-1908, driven by
// SortedGTO.C_dot_mat / mat_dot_CT / CT_dot_mat / mat_dot_C, gto/mole.py:790-938) with a different
// data model: C and C^T are passed in CSR form, and two kernels cover all four products.
#pragma once
#include "common.cuh"

// out[n][r][c] = sum_k val[k] * in[n][col[k]][c],  k in [row_ptr[r], row_ptr[r+1])
// grid (column tiles, output rows, matrices); lanes run along c: coalesced loads and stores
__global__ void spmm_left_kernel(double *out, const double *in, const int *row_ptr, const int *col,
                                 const double *val, int nrow_out, int nrow_in, int ncol)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= ncol) return;
    const double *src = in + (size_t)blockIdx.z * nrow_in * ncol;
    double s = 0.;
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) s += val[k] * src[(size_t)col[k] * ncol + c];
    out[((size_t)blockIdx.z * nrow_out + r) * ncol + c] = s;
}

// out[n][r][j] = sum_k val[k] * in[n][r][col[k]],  k in [row_ptr[j], row_ptr[j+1])
// (right multiplication by the transpose of the CSR matrix); lanes run along j: the gathers of a
// warp fall into a few neighbouring sectors of row r, the stores are coalesced
__global__ void spmm_right_kernel(double *out, const double *in, const int *row_ptr, const int *col,
                                  const double *val, int nrow, int ncol_out, int ncol_in)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= ncol_out) return;
    const double *src = in + ((size_t)blockIdx.z * nrow + r) * ncol_in;
    double s = 0.;
    for (int k = row_ptr[j]; k < row_ptr[j + 1]; ++k) s += val[k] * src[col[k]];
    out[((size_t)blockIdx.z * nrow + r) * ncol_out + j] = s;
}
