// MergerTree.h -- result records and the matching entry points, named and shaped like the
// reference's (reference src/MergerTree.h:30-51, 85-98; ShiftHalosPartsGrids src/utils.h:47).
// The bodies drive the CUDA engine through the C ABI (include/metro_b200.h); nothing is computed
// on the host.
#ifndef METRO_HOST_MERGERTREE_H
#define METRO_HOST_MERGERTREE_H
#include <cstdint>
#include <vector>

#include "Halo.h"

class MergerTree {
public:
	Halo mainHalo;
	std::vector<Halo> progHalo;
	bool isOrphan = false;
	std::vector<uint64_t> idProgenitor;
	std::vector<std::vector<int>> nCommon;     // [type][progenitor]
	void Clean();
};

void InitTrees(int nUseCat);
// FindProgenitors(0,1) runs the whole device pipeline for the pair (both directions are one
// join); FindProgenitors(1,0) then only materialises the backward trees from the same result.
void FindProgenitors(int iOne, int iTwo);
void CleanTrees(int iStep);
void FreeMergerTrees(int iStep);
void ShiftHalosPartsGrids();

#endif
