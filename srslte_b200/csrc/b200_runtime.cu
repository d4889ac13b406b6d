#include "b200_runtime.h"

#include <memory>

#include "lte_tables.h"

namespace b200 {

int DeviceArena::reserve(size_t bytes)
{
  if (bytes <= cap) {
    return B200_SUCCESS;
  }
  if (base) {
    B200_CUDA_TRY(cudaFree(base));
    base = nullptr;
    cap  = 0;
  }
  // round up so that slowly growing batches do not reallocate every call
  size_t want = (bytes + (size_t(64) << 20)) & ~((size_t(64) << 20) - 1);
  B200_CUDA_TRY(cudaMalloc(&base, want));
  cap = want;
  generation++;
  return B200_SUCCESS;
}

void* DeviceArena::take(size_t bytes)
{
  size_t off = (used + 255) & ~size_t(255);
  if (off + bytes > cap) {
    return nullptr;
  }
  used = off + bytes;
  return static_cast<char*>(base) + off;
}

void DeviceArena::release()
{
  if (base) {
    cudaFree(base);
  }
  base = nullptr;
  cap = used = 0;
  generation++;
}

int PinnedArena::reserve(size_t bytes)
{
  if (bytes <= cap) return B200_SUCCESS;
  if (base) {
    B200_CUDA_TRY(cudaFreeHost(base));
    base = nullptr;
    cap  = 0;
  }
  size_t want = (bytes + (size_t(4) << 20)) & ~((size_t(4) << 20) - 1);
  B200_CUDA_TRY(cudaHostAlloc(&base, want, cudaHostAllocDefault));
  cap = want;
  return B200_SUCCESS;
}

void* PinnedArena::take(size_t bytes)
{
  size_t off = (used + 63) & ~size_t(63);
  if (off + bytes > cap) return nullptr;
  used = off + bytes;
  return static_cast<char*>(base) + off;
}

void PinnedArena::release()
{
  if (base) cudaFreeHost(base);
  base = nullptr;
  cap = used = 0;
}

int DeviceContext::init(int dev)
{
  device = dev;
  B200_CUDA_TRY(cudaSetDevice(dev));

  std::vector<uint16_t> fwd_all, rev_all, f, r;
  qpp_off.resize(NOF_CB_SIZES);
  for (int i = 0; i < NOF_CB_SIZES; i++) {
    qpp_tables(i, f, r);
    while (fwd_all.size() % 8) { // 16-byte alignment of every table start
      fwd_all.push_back(0);
      rev_all.push_back(0);
    }
    qpp_off[i] = fwd_all.size();
    fwd_all.insert(fwd_all.end(), f.begin(), f.end());
    rev_all.insert(rev_all.end(), r.begin(), r.end());
  }
  B200_CUDA_TRY(cudaMalloc(&qpp_fwd_all, fwd_all.size() * sizeof(uint16_t)));
  B200_CUDA_TRY(cudaMalloc(&qpp_rev_all, rev_all.size() * sizeof(uint16_t)));
  B200_CUDA_TRY(cudaMemcpy(qpp_fwd_all, fwd_all.data(), fwd_all.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemcpy(qpp_rev_all, rev_all.data(), rev_all.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));

  return B200_SUCCESS;
}

int DeviceContext::crc_visit(int cb_idx, int kind, const CrcPow** nat, const CrcPow** perm)
{
  std::lock_guard<std::mutex> lock(crc_mutex);
  RmTableKey                  key{cb_idx, kind};
  auto                        it = crc_tables.find(key);
  if (it == crc_tables.end()) {
    std::vector<CrcPow> n, p;
    crc_visit_tables(kind == 0 ? CRC24A_POLY : CRC24B_POLY, cb_idx, n, p);
    CrcTables t;
    B200_CUDA_TRY(cudaSetDevice(device));
    B200_CUDA_TRY(cudaMalloc(&t.nat, n.size() * sizeof(CrcPow)));
    B200_CUDA_TRY(cudaMalloc(&t.perm, p.size() * sizeof(CrcPow)));
    B200_CUDA_TRY(cudaMemcpy(t.nat, n.data(), n.size() * sizeof(CrcPow), cudaMemcpyHostToDevice));
    B200_CUDA_TRY(cudaMemcpy(t.perm, p.data(), p.size() * sizeof(CrcPow), cudaMemcpyHostToDevice));
    it = crc_tables.emplace(key, t).first;
  }
  *nat  = it->second.nat;
  *perm = it->second.perm;
  return B200_SUCCESS;
}

const uint16_t* DeviceContext::rm_table(int cb_idx, int rv)
{
  std::lock_guard<std::mutex> lock(rm_mutex);
  RmTableKey                  key{cb_idx, rv};
  auto                        it = rm_gather.find(key);
  if (it != rm_gather.end()) {
    return it->second;
  }
  std::vector<uint16_t> inv;
  rm_gather_table(cb_idx, rv, inv);
  uint16_t* d = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&d, inv.size() * sizeof(uint16_t)) != cudaSuccess ||
      cudaMemcpy(d, inv.data(), inv.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    B200_LOG_ERROR("could not upload rate-matching table cb_idx=%d rv=%d", cb_idx, rv);
    return nullptr;
  }
  rm_gather[key] = d;
  return d;
}

DeviceContext* device_context(int device)
{
  static std::mutex                                    m;
  static std::map<int, std::unique_ptr<DeviceContext>> ctxs;
  std::lock_guard<std::mutex>                          lock(m);
  auto                                                 it = ctxs.find(device);
  if (it != ctxs.end()) {
    return it->second.get();
  }
  std::unique_ptr<DeviceContext> c(new DeviceContext());
  if (c->init(device) != B200_SUCCESS) {
    return nullptr;
  }
  DeviceContext* raw = c.get();
  ctxs[device]       = std::move(c);
  return raw;
}

} // namespace b200
