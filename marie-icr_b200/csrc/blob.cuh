// Flat weight blob shared by mb_load_craft / mb_load_trocr (produced by marie-icr_b200/weights.py).
//   header : "MB2W" | u32 version | u32 n_entries | u32 reserved
//   entry  : char name[64] | u32 dtype (0 f32, 1 bf16, 2 i32) | u32 ndim | u64 dims[4] | u64 offset | u64 nbytes
//   data   : tensors at 256-byte aligned offsets relative to the start of the data section
#pragma once
#include "common.cuh"
#include <map>

struct BlobTensor {
    void* dev = nullptr;
    int dtype = 0;
    int ndim = 0;
    long long dims[4] = {0, 0, 0, 0};
    size_t nbytes = 0;
};

struct WeightBlob {
    void* dev_base = nullptr;
    size_t dev_bytes = 0;
    std::map<std::string, BlobTensor> tensors;

    int load(mb_ctx* ctx, const void* host, size_t nbytes) {
        const unsigned char* p = (const unsigned char*)host;
        if (nbytes < 16 || memcmp(p, "MB2W", 4) != 0) return mb_set_err(ctx, MB_ERR_ARG, "weight blob: bad magic");
        uint32_t version, n;
        memcpy(&version, p + 4, 4);
        memcpy(&n, p + 8, 4);
        if (version != 1) return mb_set_err(ctx, MB_ERR_ARG, "weight blob: unsupported version %u", version);
        const size_t entry_bytes = 64 + 4 + 4 + 32 + 8 + 8;
        const size_t header = 16 + (size_t)n * entry_bytes;
        const size_t data_off = mb_align_up(header, 256);
        if (nbytes < data_off) return mb_set_err(ctx, MB_ERR_ARG, "weight blob: truncated");
        dev_bytes = nbytes - data_off;
        if (cudaMalloc(&dev_base, dev_bytes ? dev_bytes : 256) != cudaSuccess) {
            cudaGetLastError();
            return mb_set_err(ctx, MB_ERR_OOM, "weight blob: cudaMalloc(%zu) failed", dev_bytes);
        }
        MB_CUDA(ctx, cudaMemcpy(dev_base, p + data_off, dev_bytes, cudaMemcpyHostToDevice));
        for (uint32_t i = 0; i < n; ++i) {
            const unsigned char* e = p + 16 + (size_t)i * entry_bytes;
            char name[65];
            memcpy(name, e, 64);
            name[64] = 0;
            BlobTensor t;
            uint32_t dtype, ndim;
            memcpy(&dtype, e + 64, 4);
            memcpy(&ndim, e + 68, 4);
            uint64_t dims[4], off, nb;
            memcpy(dims, e + 72, 32);
            memcpy(&off, e + 104, 8);
            memcpy(&nb, e + 112, 8);
            if (off + nb > dev_bytes) return mb_set_err(ctx, MB_ERR_ARG, "weight blob: entry %s out of range", name);
            t.dev = (unsigned char*)dev_base + off;
            t.dtype = (int)dtype;
            t.ndim = (int)ndim;
            for (int d = 0; d < 4; ++d) t.dims[d] = (long long)dims[d];
            t.nbytes = nb;
            tensors[name] = t;
        }
        return 0;
    }
    const BlobTensor* get(const std::string& name) const {
        auto it = tensors.find(name);
        return it == tensors.end() ? nullptr : &it->second;
    }
    void release() {
        if (dev_base) cudaFree(dev_base);
        dev_base = nullptr;
        tensors.clear();
    }
};
