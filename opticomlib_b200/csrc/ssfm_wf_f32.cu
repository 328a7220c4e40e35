// k_wf launch side, float instantiations (see ssfm_wf_impl.inl)
#define SSFM_WF_REAL float
#include "ssfm_wf_impl.inl"
