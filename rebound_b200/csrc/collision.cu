// collision.cu -- placeholder, replaced below in this round.
#include "engine.cuh"
int collision_search(rebcu_handle* h, const rebcu_config* c) { (void)c; return rebcu_fail(h, REBCU_ERR_ARG, "collision search not built yet"); }
