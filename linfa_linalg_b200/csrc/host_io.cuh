// Host views <-> the engine's column-major HBM layout: packing of arbitrary-stride ndarray-style views
// (tests/common.rs:12-43 randomises them) on upload, scatter back into the caller's storage on download.
// Shared by the single-device entry points (api.cu) and the multi-device ones (multi.cu).
#pragma once
#include "common.cuh"

namespace lfb {

enum Layout { L_ROW, L_COL, L_GEN };

inline Layout classify(int64_t rows, int64_t cols, int64_t rs, int64_t cs, int64_t *ld) {
    if (cs == 1 && (rows == 1 || rs >= cols) && rs > 0) { *ld = rows == 1 ? cols : rs; return L_ROW; }
    if (rows == 1 && cs == 1) { *ld = cols; return L_ROW; }
    if (rs == 1 && (cols == 1 || cs >= rows) && cs > 0) { *ld = cols == 1 ? rows : cs; return L_COL; }
    if (cols == 1 && rs == 1) { *ld = rows; return L_COL; }
    return L_GEN;
}

// Host view -> device column-major (rows x cols, ldd).
template <typename T>
void upload(lfb_handle &h, const T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *d, int64_t ldd) {
    if (rows <= 0 || cols <= 0) return;
    int64_t ld = 0;
    Layout lay = classify(rows, cols, rs, cs, &ld);
    if (lay == L_COL) {
        LFB_CUDA(cudaMemcpy2DAsync(d, ldd * sizeof(T), a, ld * sizeof(T), rows * sizeof(T), cols, cudaMemcpyHostToDevice, h.stream));
        return;
    }
    const T *src = a;
    if (lay == L_GEN) {
        T *pk = (T *)h.pinned_buf(sizeof(T) * rows * cols);
        for (int64_t i = 0; i < rows; ++i)
            for (int64_t j = 0; j < cols; ++j) pk[i * cols + j] = a[i * rs + j * cs];
        src = pk;
        ld = cols;
    }
    DevBuf<T> tmp(h, (size_t)ld * rows);
    LFB_CUDA(cudaMemcpyAsync(tmp.get(), src, sizeof(T) * ((rows - 1) * ld + cols), cudaMemcpyHostToDevice, h.stream));
    transpose<T>(h, tmp.get(), cols, rows, ld, d, ldd);
    if (lay == L_GEN) LFB_CUDA(cudaStreamSynchronize(h.stream));  // pinned staging buffer is reused
}

// Device column-major -> host view (in place into the caller's storage).  Synchronises.
template <typename T>
void download(lfb_handle &h, const T *d, int64_t ldd, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs) {
    if (rows <= 0 || cols <= 0) return;
    int64_t ld = 0;
    Layout lay = classify(rows, cols, rs, cs, &ld);
    if (lay == L_COL) {
        LFB_CUDA(cudaMemcpy2DAsync(a, ld * sizeof(T), d, ldd * sizeof(T), rows * sizeof(T), cols, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        return;
    }
    if (lay == L_GEN) ld = cols;
    DevBuf<T> tmp(h, (size_t)ld * rows);
    transpose<T>(h, d, rows, cols, ldd, tmp.get(), ld);   // tmp: cols x rows column-major == row-major rows x cols
    if (lay == L_ROW) {
        // copy row by row extents only (do not touch padding between rows)
        LFB_CUDA(cudaMemcpy2DAsync(a, ld * sizeof(T), tmp.get(), ld * sizeof(T), cols * sizeof(T), rows, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaStreamSynchronize(h.stream));
    } else {
        T *pk = (T *)h.pinned_buf(sizeof(T) * rows * cols);
        LFB_CUDA(cudaMemcpyAsync(pk, tmp.get(), sizeof(T) * rows * cols, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        for (int64_t i = 0; i < rows; ++i)
            for (int64_t j = 0; j < cols; ++j) a[i * rs + j * cs] = pk[i * cols + j];
    }
}

template <typename T>
void upload_vec(lfb_handle &h, const T *v, int64_t n, T *d) {
    if (n > 0) LFB_CUDA(cudaMemcpyAsync(d, v, sizeof(T) * n, cudaMemcpyHostToDevice, h.stream));
}
template <typename T>
void download_vec(lfb_handle &h, const T *d, int64_t n, T *v) {
    if (n > 0) LFB_CUDA(cudaMemcpyAsync(v, d, sizeof(T) * n, cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaStreamSynchronize(h.stream));
}


}  // namespace lfb
