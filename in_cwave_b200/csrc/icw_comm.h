// icw_comm.h -- NCCL plumbing of the time-sharded path (icw_comm.cu); NCCL is dlopen'ed, never linked.
#pragma once
#include <cstddef>
#include <string>
#include <cuda_runtime.h>

#define ICW_COMM_ID_BYTES 128

namespace icw {

bool nccl_load();
const char *nccl_why();
int nccl_version();
int comm_unique_id(unsigned char out[ICW_COMM_ID_BYTES], std::string &err);
int comm_init(const unsigned char id[ICW_COMM_ID_BYTES], int rank, int world, void **comm, std::string &err);
int comm_destroy(void *comm, std::string &err);
int comm_shift_right(void *comm, int rank, int world, const double *d_send, double *d_recv, size_t n_doubles,
                     cudaStream_t st, std::string &err);
int comm_reduce(void *comm, unsigned long long *d_sum, size_t n_sum, double *d_max, size_t n_max, cudaStream_t st, std::string &err);

}  // namespace icw
