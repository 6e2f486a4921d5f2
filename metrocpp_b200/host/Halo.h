// Halo.h -- host-side catalogue record, same public fields as the reference's Halo class
// (reference src/Halo.h:30-71) so that code written against it keeps compiling.
#ifndef METRO_HOST_HALO_H
#define METRO_HOST_HALO_H
#include <cstdint>

#ifndef NPTYPES_MAX
#define NPTYPES_MAX 8
#endif

class Halo {
public:
	Halo();

	float mTot, mDM, mGas, mFM, mStar, rVir;
	float cNFW, rsNFW;
	float lambda, vMax, sigV;
	float X[3], V[3], L[3];
	float fMhires;

	bool isToken;              // token halo standing in for an orphan (reference src/Halo.h:43)
	int nOrphanSteps;          // steps spent without a progenitor
	int nSub;
	int nPart[NPTYPES_MAX];    // per particle type; only the first nPTypes entries are used
	uint64_t ID, hostID;

	int nAllPart() const;      // sum over the particle types (reference src/Halo.cpp:63-70)
	float Distance(const float *pos) const;
};

#endif
