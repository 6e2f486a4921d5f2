// global_vars.h -- the process-wide state the reference keeps in globals (reference
// src/global_vars.h:35-130), reduced to what the matching path and its I/O touch.  The particle
// containers (locParts / locMapParts, std::map<pid, vector<Particle>>) are replaced by flat
// structure-of-arrays tuples: that is what the device ingests.
#ifndef METRO_HOST_GLOBAL_VARS_H
#define METRO_HOST_GLOBAL_VARS_H
#include <cstdint>
#include <string>
#include <memory>
#include <utility>
#include <vector>

#include "../../include/metro_b200.h"
#include "Halo.h"
#include "MergerTree.h"

// std::allocator that leaves trivially constructible elements uninitialised on resize(): the ingest fills every element
// itself (in parallel), a serial zero-fill of the three tuple columns first would cost as much as the parse
template <typename T>
struct UninitAlloc : std::allocator<T> {
	template <typename U> struct rebind { using other = UninitAlloc<U>; };
	UninitAlloc() = default;
	template <typename U> UninitAlloc(const UninitAlloc<U> &) {}
	template <typename U, typename... Args> void construct(U *p, Args &&...args)
	{
		if constexpr (sizeof...(Args) == 0) ::new ((void *) p) U;
		else ::new ((void *) p) U(std::forward<Args>(args)...);
	}
};

struct TupleSoA {                      // one catalogue's "lines of the _particles files"
	std::vector<uint64_t, UninitAlloc<uint64_t>> pid;
	std::vector<uint32_t, UninitAlloc<uint32_t>> halo;    // index into locHalos[i]
	std::vector<uint8_t, UninitAlloc<uint8_t>> type;
	std::vector<int64_t> haloOffset;   // [nHalos + 1] when the tuples are grouped by halo in catalogue order (else empty)
	void clear()
	{
		pid.clear(); halo.clear(); type.clear(); haloOffset.clear();
		pid.shrink_to_fit(); halo.shrink_to_fit(); type.shrink_to_fit(); haloOffset.shrink_to_fit();
	}
};

extern int locTask, totTask;                                // rank / ranks (one GPU each), reference src/global_vars.h:38
extern int gpuDevice;                                        // CUDA device of this rank
extern uint64_t pidShardWidth;                               // multi-rank: rank r keeps the particle IDs [r * width, (r+1) * width) (last rank: and above)
extern std::vector<std::vector<Halo>> locHalos;             // [2][nHalos]
extern std::vector<TupleSoA> locTuples;                     // [2]
extern std::vector<std::vector<MergerTree>> locMTrees;      // [2]
extern std::vector<std::vector<MergerTree>> locCleanTrees;  // [step]
extern std::vector<uint64_t> allOrphIDs;
extern int nLocHalos[2];
extern int nPTypes;
extern bool noPType;                                        // -DNOPTYPE behaviour: every type = 1
extern int treeFlavour;                                     // METRO_FLAVOUR_*
extern int minPartCmp, minPartHalo, facOrphanSteps, maxOrphanSteps;
extern int nSnapsUse, nSnaps, nChunks, nGrid, nTreeChunks;
extern int iNumCat, iUseCat;
extern float boxSize;
extern std::string cosmologicalModel;
extern metro_ctx *gpuCtx;                                   // the engine context (one per process)
extern bool keepRawTrees;                                   // fill locMTrees[] after FindProgenitors (off: only clean trees)
extern bool slotDirty[2];                                   // host catalogue changed since last upload

void InitLocVariables(bool needGpu = true);
void CleanMemory(int iCat);
metro_params CurrentParams();
[[noreturn]] void Fatal(const std::string &msg);            // prints and exits non-zero

#endif
