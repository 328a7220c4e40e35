// k_wf launch side, double instantiations (see ssfm_wf_impl.inl)
#define SSFM_WF_REAL double
#include "ssfm_wf_impl.inl"
