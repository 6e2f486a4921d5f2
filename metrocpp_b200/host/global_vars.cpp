#include "global_vars.h"

#include <cstdio>
#include <cstdlib>

int locTask = 0, totTask = 1;
int gpuDevice = 0;
uint64_t pidShardWidth = 0;
std::vector<std::vector<Halo>> locHalos;
std::vector<TupleSoA> locTuples;
std::vector<std::vector<MergerTree>> locMTrees;
std::vector<std::vector<MergerTree>> locCleanTrees;
std::vector<uint64_t> allOrphIDs;
int nLocHalos[2] = {0, 0};
int nPTypes = 2;
bool noPType = true;
int treeFlavour = METRO_FLAVOUR_ZOOM_DUP;
int minPartCmp = 0, minPartHalo = 0, facOrphanSteps = 0, maxOrphanSteps = 0;
int nSnapsUse = 0, nSnaps = 0, nChunks = 1, nGrid = 0, nTreeChunks = 1;
int iNumCat = 0, iUseCat = 0;
float boxSize = 0.0f;
std::string cosmologicalModel;
metro_ctx *gpuCtx = nullptr;
bool keepRawTrees = false;
bool slotDirty[2] = {true, true};

void InitLocVariables(bool needGpu)
{
	locHalos.assign(2, {});
	locTuples.assign(2, {});
	locMTrees.assign(2, {});
	if (needGpu && !gpuCtx) {
		int rc = metro_ctx_create(gpuDevice, &gpuCtx);
		if (rc != METRO_OK) Fatal(std::string("cannot create the GPU context: ") + metro_last_error(nullptr));
	}
}

void CleanMemory(int iCat)
{
	locHalos[iCat].clear();
	locHalos[iCat].shrink_to_fit();
	locTuples[iCat].clear();
	nLocHalos[iCat] = 0;
	slotDirty[iCat] = true;
}

metro_params CurrentParams()
{
	metro_params p;
	p.nptypes = nPTypes;
	p.min_part_cmp = minPartCmp;
	p.min_part_halo = minPartHalo;
	p.fac_orphan_steps = facOrphanSteps;
	p.max_orphan_steps = maxOrphanSteps;
	p.flavour = treeFlavour;
	return p;
}

void Fatal(const std::string &msg)
{
	// the reference prints and exit(0)s (src/utils.cpp:74-78); a failed run must not look successful
	fprintf(stderr, "MetroCPP-b200: %s\n", msg.c_str());
	exit(2);
}
