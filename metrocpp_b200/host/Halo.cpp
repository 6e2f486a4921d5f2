#include "Halo.h"

#include <cmath>

#include "global_vars.h"

Halo::Halo()
{
	mTot = mDM = mGas = mFM = mStar = rVir = 0.0f;
	cNFW = rsNFW = lambda = vMax = sigV = 0.0f;
	fMhires = 1.0f;
	for (int i = 0; i < 3; i++) { X[i] = -1.0f; V[i] = 0.0f; L[i] = 0.0f; }
	isToken = false;
	nOrphanSteps = 0;
	nSub = 0;
	for (int t = 0; t < NPTYPES_MAX; t++) nPart[t] = 0;   // the reference leaves most of these unset
	ID = hostID = 0;
}

int Halo::nAllPart() const
{
	int n = 0;
	for (int t = 0; t < nPTypes; t++) n += nPart[t];
	return n;
}

float Halo::Distance(const float *pos) const
{
	float d2 = 0.0f;
	for (int i = 0; i < 3; i++) d2 += (pos[i] - X[i]) * (pos[i] - X[i]);
	return std::sqrt(d2);
}
