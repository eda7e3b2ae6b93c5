// comm.h — NCCL, bound at run time (dlopen) so that libqlcuda.so has no link-time dependency on it: a single-GPU host never
// loads it, a multi-GPU host gets whichever libnccl.so.2 its process already uses (e.g. the one PyTorch ships) or the system one.
// Only use: the episode-statistics reduction of the env shards — nothing on the step path.
#pragma once
#include <cstddef>
#include <string>

namespace qlc_comm {

// the handful of NCCL declarations used (ABI-stable since NCCL 2.0; values from nccl.h)
struct UniqueId { char internal[128]; };
typedef struct ncclComm* Comm;
enum { kFloat64 = 8 };
enum { kSum = 0, kMax = 2 };

struct Api {
    int (*GetVersion)(int*);
    int (*GetUniqueId)(UniqueId*);
    int (*CommInitRank)(Comm*, int, UniqueId, int);
    int (*CommInitRankConfig)(Comm*, int, UniqueId, int, void* /*ncclConfig_t*/);   // optional (NCCL >= 2.14): NULL if the library lacks it
    int (*CommDestroy)(Comm);
    int (*CommCount)(Comm, int*);
    int (*AllGather)(const void*, void*, size_t, int, Comm, void* /*cudaStream_t*/);
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, void* /*cudaStream_t*/);
    const char* (*GetErrorString)(int);
    std::string path;      // what was loaded
};

// ncclConfig_t as of NCCL 2.18 (nccl.h: size, magic, version, then the options; NCCL accepts older / shorter layouts by `size` and
// `version` and fills what is missing with its defaults). Only use: min/maxCTAs = 1 - a 40-byte all-gather needs one CTA, and every
// SM slot the collective holds is one the cooperative step launch has to wait for.
struct Config {
    size_t size; unsigned int magic; unsigned int version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs;
    const char* netName;
    int splitShare;
};
inline Config one_cta_config() {
    const int undef = (int)0x80000000;      // NCCL_CONFIG_UNDEF_INT
    return Config{sizeof(Config), 0xcafebeefu, 21800u, undef, undef, 1, 1, nullptr, undef};
}

// NULL + *why when libnccl cannot be loaded (QLC_NCCL_LIB overrides the name). Thread-safe, loaded once.
const Api* api(std::string* why);

}  // namespace qlc_comm
