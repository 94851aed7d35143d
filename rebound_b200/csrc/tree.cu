// tree.cu -- placeholder, replaced below in this round.
#include "engine.cuh"
void tree_free(rebcu_handle* h) { (void)h; }
int tree_build(rebcu_handle* h, const rebcu_config* c) { (void)c; return rebcu_fail(h, REBCU_ERR_ARG, "tree not built yet"); }
int tree_gravity(rebcu_handle* h, rebcu_config* c) { (void)c; return rebcu_fail(h, REBCU_ERR_ARG, "tree not built yet"); }
