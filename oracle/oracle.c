/*
 * oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's many-particle hot path
 * (hannorein/rebound v5.0.0; file:line citations are relative to the reference root).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (rebound_b200/csrc) never links or calls it.
 *
 * Parity pin: every function here is checked bit-for-bit against the unmodified reference
 * (oracle/_ref/libref_harness*.so, built by oracle/Makefile from /root/reference/src) in
 * tests/test_oracle_vs_reference.py, and against the committed fixtures in tests/golden/
 * (generated from the reference by tests/golden/make_golden.py) in tests/test_oracle_golden.py.
 *
 * The restatement is written in "gather" form -- every particle accumulates its own sum over
 * sources in ascending index order -- which is what the reference's OpenMP build does
 * (src/gravity.c:216-232, 309-414) and what its serial build produces bit-for-bit when no ghost
 * boxes are in use (determinism facts).  Compile with -ffp-contract=off.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "../include/rebound_b200.h"

static char orc_errbuf[512];
const char* orc_last_error(void){ return orc_errbuf; }
static int orc_fail(int code, const char* msg){
    strncpy(orc_errbuf, msg, sizeof(orc_errbuf)-1);
    return code;
}

/* ------------------------------------------------------------------------------------------ */
/* Ghost boxes: src/boundary.c:145-201                                                         */
/* ------------------------------------------------------------------------------------------ */
static rebcu_vec6d ghostbox(const rebcu_config* c, int i, int j, int k){
    const double bx = c->root_size*(double)c->N_root_x;
    const double by = c->root_size*(double)c->N_root_y;
    const double bz = c->root_size*(double)c->N_root_z;
    rebcu_vec6d gb = {0,0,0,0,0,0};
    if (c->boundary==REBCU_BOUNDARY_OPEN || c->boundary==REBCU_BOUNDARY_PERIODIC){
        gb.x = bx*(double)i; gb.y = by*(double)j; gb.z = bz*(double)k;
    }else if (c->boundary==REBCU_BOUNDARY_SHEAR){
        gb.vy = -1.5*(double)i*c->OMEGA*bx;
        double shift;
        if (i==0)      shift = -fmod(gb.vy*c->t, by);
        else if (i>0)  shift = -fmod(gb.vy*c->t - by/2., by) - by/2.;
        else           shift = -fmod(gb.vy*c->t + by/2., by) + by/2.;
        gb.x = bx*(double)i; gb.y = by*(double)j - shift; gb.z = bz*(double)k;
    }
    return gb;
}

/* Terms excluded by gravity_ignore_terms: src/gravity.c:219-221, 263-264, 312-314. */
static int ignored(int terms, uint64_t i, uint64_t j){
    if (i==j) return 1;
    if (terms==REBCU_IGNORE_TERMS_BETWEEN_0_AND_1 && ((j==1 && i==0) || (i==1 && j==0))) return 1;
    if (terms==REBCU_IGNORE_TERMS_INVOLVING_0 && (j==0 || i==0)) return 1;
    return 0;
}

/* Sources of particle i: the active particles, plus the test particles if i is active and
 * testparticle_type==1 (src/gravity.c:217-218, 259-262). */
static uint64_t source_count(const rebcu_config* c, uint64_t N, uint64_t i){
    const uint64_t Na = (c->N_active==REBCU_SIZE_MAX)?N:c->N_active;
    return (i<Na && c->testparticle_type) ? N : Na;
}

/* ------------------------------------------------------------------------------------------ */
/* Direct summation: src/gravity.c:167-282 (BASIC) and :284-531 (COMPENSATED)                  */
/* ------------------------------------------------------------------------------------------ */
static void gravity_basic(const rebcu_config* c, rebcu_particle* p, uint64_t N){
    const double G = c->G;
    const double soft2 = c->softening*c->softening;
    const int ngb = (2*c->N_ghost_x+1)*(2*c->N_ghost_y+1)*(2*c->N_ghost_z+1);
    rebcu_vec6d* gbs = malloc(sizeof(rebcu_vec6d)*ngb);
    int n=0;
    for (int gx=-c->N_ghost_x; gx<=c->N_ghost_x; gx++)
        for (int gy=-c->N_ghost_y; gy<=c->N_ghost_y; gy++)
            for (int gz=-c->N_ghost_z; gz<=c->N_ghost_z; gz++)
                gbs[n++] = ghostbox(c, gx, gy, gz);
#pragma omp parallel for schedule(dynamic,64)
    for (uint64_t i=0;i<N;i++){
        double ax=0., ay=0., az=0.;
        const uint64_t ns = source_count(c, N, i);
        for (int g=0; g<ngb; g++){
            const double xi = gbs[g].x + p[i].x;
            const double yi = gbs[g].y + p[i].y;
            const double zi = gbs[g].z + p[i].z;
            for (uint64_t j=0;j<ns;j++){
                if (ignored(c->gravity_ignore_terms, i, j)) continue;
                const double dx = xi - p[j].x;
                const double dy = yi - p[j].y;
                const double dz = zi - p[j].z;
                const double rr = sqrt(dx*dx + dy*dy + dz*dz + soft2);
                const double pre = -G/(rr*rr*rr)*p[j].m;
                ax += pre*dx; ay += pre*dy; az += pre*dz;
            }
        }
        p[i].ax = ax; p[i].ay = ay; p[i].az = az;
    }
    free(gbs);
}

static void gravity_compensated(const rebcu_config* c, rebcu_particle* p, uint64_t N, double* cs_out){
    const double G = c->G;
    const double soft2 = c->softening*c->softening;
#pragma omp parallel for schedule(dynamic,64)
    for (uint64_t i=0;i<N;i++){
        double s[3] = {0.,0.,0.};   /* running sums   (particles[i].ax..az) */
        double e[3] = {0.,0.,0.};   /* compensation   (r->gravity_cs[i])    */
        const uint64_t ns = source_count(c, N, i);
        for (uint64_t j=0;j<ns;j++){
            if (ignored(c->gravity_ignore_terms, i, j)) continue;
            double d[3] = { p[i].x - p[j].x, p[i].y - p[j].y, p[i].z - p[j].z };
            const double r2 = d[0]*d[0] + d[1]*d[1] + d[2]*d[2] + soft2;
            const double rr = sqrt(r2);
            const double pre = G/(r2*rr);
            const double prej = -pre*p[j].m;
            for (int k=0;k<3;k++){          /* Kahan step, gravity.c:323-341 */
                const double term = prej*d[k];
                const double y = term - e[k];
                const double t = s[k] + y;
                e[k] = (t - s[k]) - y;
                s[k] = t;
            }
        }
        p[i].ax = s[0]; p[i].ay = s[1]; p[i].az = s[2];
        if (cs_out){ cs_out[3*i] = e[0]; cs_out[3*i+1] = e[1]; cs_out[3*i+2] = e[2]; }   /* r->gravity_cs[i] */
    }
}

/* Selected rows of the direct sum (full-size spot checks: one row of N = 2^22 is 4e6 pair terms): the accelerations of
 * particles rows[0..n_rows) from all their sources, by the loop bodies above (tests/test_oracle_vs_reference.py checks
 * that the rows equal the full evaluation bit for bit).  out = ax,ay,az per row.  BASIC without ghost boxes or COMPENSATED. */
int orc_gravity_rows(rebcu_config* c, rebcu_particle* p, uint64_t N, const uint64_t* rows, uint64_t n_rows, double* out){
    const double G = c->G;
    const double soft2 = c->softening*c->softening;
    if (c->gravity!=REBCU_GRAVITY_BASIC && c->gravity!=REBCU_GRAVITY_COMPENSATED) return orc_fail(REBCU_ERR_ARG, "orc_gravity_rows: direct summation only");
    if (c->N_ghost_x || c->N_ghost_y || c->N_ghost_z) return orc_fail(REBCU_ERR_ARG, "orc_gravity_rows: no ghost boxes");
    const rebcu_vec6d gb0 = ghostbox(c, 0, 0, 0);
#pragma omp parallel for schedule(dynamic,1)
    for (uint64_t k=0;k<n_rows;k++){
        const uint64_t i = rows[k];
        const uint64_t ns = source_count(c, N, i);
        if (c->gravity==REBCU_GRAVITY_BASIC){
            double ax=0., ay=0., az=0.;
            const double xi = gb0.x + p[i].x, yi = gb0.y + p[i].y, zi = gb0.z + p[i].z;
            for (uint64_t j=0;j<ns;j++){
                if (ignored(c->gravity_ignore_terms, i, j)) continue;
                const double dx = xi - p[j].x;
                const double dy = yi - p[j].y;
                const double dz = zi - p[j].z;
                const double rr = sqrt(dx*dx + dy*dy + dz*dz + soft2);
                const double pre = -G/(rr*rr*rr)*p[j].m;
                ax += pre*dx; ay += pre*dy; az += pre*dz;
            }
            out[3*k] = ax; out[3*k+1] = ay; out[3*k+2] = az;
        }else{
            double s[3] = {0.,0.,0.}, e[3] = {0.,0.,0.};
            for (uint64_t j=0;j<ns;j++){
                if (ignored(c->gravity_ignore_terms, i, j)) continue;
                double d[3] = { p[i].x - p[j].x, p[i].y - p[j].y, p[i].z - p[j].z };
                const double r2 = d[0]*d[0] + d[1]*d[1] + d[2]*d[2] + soft2;
                const double rr = sqrt(r2);
                const double pre = G/(r2*rr);
                const double prej = -pre*p[j].m;
                for (int q=0;q<3;q++){
                    const double term = prej*d[q];
                    const double y = term - e[q];
                    const double t = s[q] + y;
                    e[q] = (t - s[q]) - y;
                    s[q] = t;
                }
            }
            out[3*k] = s[0]; out[3*k+1] = s[1]; out[3*k+2] = s[2];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Jerk kick of the modified-kick schemes (EOS): src/gravity.c:850-924                          */
/* ------------------------------------------------------------------------------------------ */
/* One pair (i > j): velocity increments from the gradient of |a_i - a_j| terms.  `both`: also update j. */
static void jerk_pair(rebcu_particle* p, uint64_t i, uint64_t j, double vG2, int both){
    const double d[3]  = { p[i].x - p[j].x,   p[i].y - p[j].y,   p[i].z - p[j].z };
    const double da[3] = { p[i].ax - p[j].ax, p[i].ay - p[j].ay, p[i].az - p[j].az };
    const double dr = sqrt(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]);
    const double alphasum = da[0]*d[0] + da[1]*d[1] + da[2]*d[2];
    const double pf2 = vG2/(dr*dr*dr);                     /* gravity.c:876 */
    const double pf1 = alphasum*pf2/dr*3./dr;              /* gravity.c:879 */
    const double pf1i = pf1*p[j].m, pf2i = pf2*p[j].m;
    double* vi = &p[i].vx; double* vj = &p[j].vx;
    for (int k=0;k<3;k++) vi[k] += d[k]*pf1i - da[k]*pf2i;          /* gravity.c:882-884, 909-911 */
    if (both){
        const double pf1j = pf1*p[i].m, pf2j = pf2*p[i].m;
        for (int k=0;k<3;k++) vj[k] += da[k]*pf2j - d[k]*pf1j;      /* gravity.c:885-887, 915-917 */
    }
}

static void apply_jerk(const rebcu_config* c, rebcu_particle* p, uint64_t N, double v){
    const uint64_t Na = (c->N_active==REBCU_SIZE_MAX)?N:(c->N_active<N?c->N_active:N);
    const uint64_t starti = (c->gravity_ignore_terms==REBCU_IGNORE_TERMS_NONE)?1:2;         /* gravity.c:857 */
    const uint64_t startj = (c->gravity_ignore_terms==REBCU_IGNORE_TERMS_INVOLVING_0)?1:0;  /* gravity.c:858 */
    const double vG2 = 2.*v*c->G;
    /* massive-massive pairs (gravity.c:861-889), serial order: the scatter makes the order part of the result */
    for (uint64_t i=starti;i<Na;i++)
        for (uint64_t j=startj;j<i;j++) jerk_pair(p, i, j, vG2, 1);
    /* rows of the test particles: every j < i, not only the massive ones (gravity.c:892-920); the back reaction
     * only with testparticle_type != 0 */
    for (uint64_t i=Na;i<N;i++)
        for (uint64_t j=startj;j<i;j++) jerk_pair(p, i, j, vG2, c->testparticle_type!=0);
}

/* The same result as a GATHER (what the CUDA kernel does; kept here so that the order analysis is checked on the
 * CPU): in the serial scatter above, the velocity of particle k first receives its own row (partners p < k, as the
 * i-side of the pair, ascending p) and then, as i runs on, the back reaction of every later row p > k (as the
 * j-side, ascending p) -- i.e. one sum over ascending p.  Which pairs exist:
 *   p < k: (i=k, j=p) needs p >= startj, and k >= starti when k is massive (test-particle rows have no starti)
 *   p > k: (i=p, j=k) needs k >= startj, and p >= starti when p is massive, testparticle_type != 0 otherwise. */
int orc_apply_jerk_gather(rebcu_config* c, rebcu_particle* p, uint64_t N, double v){
    const uint64_t Na = (c->N_active==REBCU_SIZE_MAX)?N:(c->N_active<N?c->N_active:N);
    const uint64_t starti = (c->gravity_ignore_terms==REBCU_IGNORE_TERMS_NONE)?1:2;
    const uint64_t startj = (c->gravity_ignore_terms==REBCU_IGNORE_TERMS_INVOLVING_0)?1:0;
    const double vG2 = 2.*v*c->G;
    double* out = malloc(sizeof(double)*3*(N?N:1));
    for (uint64_t k=0;k<N;k++){
        double w[3] = { p[k].vx, p[k].vy, p[k].vz };
        for (uint64_t q=startj;q<N;q++){
            if (q==k) continue;
            const int k_is_i = q<k;
            if (k_is_i){ if (k<Na && k<starti) continue; }
            else { if (k<startj) continue; if (q<Na ? q<starti : c->testparticle_type==0) continue; }
            const uint64_t i = k_is_i ? k : q, j = k_is_i ? q : k;
            const double d[3]  = { p[i].x - p[j].x,   p[i].y - p[j].y,   p[i].z - p[j].z };
            const double da[3] = { p[i].ax - p[j].ax, p[i].ay - p[j].ay, p[i].az - p[j].az };
            const double dr = sqrt(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]);
            const double alphasum = da[0]*d[0] + da[1]*d[1] + da[2]*d[2];
            const double pf2 = vG2/(dr*dr*dr);
            const double pf1 = alphasum*pf2/dr*3./dr;
            const double pf1o = pf1*p[q].m, pf2o = pf2*p[q].m;          /* scaled with the OTHER particle's mass */
            for (int a=0;a<3;a++) w[a] += k_is_i ? (d[a]*pf1o - da[a]*pf2o) : (da[a]*pf2o - d[a]*pf1o);
        }
        out[3*k]=w[0]; out[3*k+1]=w[1]; out[3*k+2]=w[2];
    }
    for (uint64_t k=0;k<N;k++){ p[k].vx=out[3*k]; p[k].vy=out[3*k+1]; p[k].vz=out[3*k+2]; }
    free(out);
    return 0;
}

int orc_apply_jerk(rebcu_config* c, rebcu_particle* p, uint64_t N, double v){
    orc_errbuf[0]=0;
    apply_jerk(c, p, N, v);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Boundary check: src/boundary.c:35-141                                                       */
/* ------------------------------------------------------------------------------------------ */
static void boundary_check(rebcu_config* c, rebcu_particle* p, uint64_t* Np){
    const double bx = c->root_size*(double)c->N_root_x;
    const double by = c->root_size*(double)c->N_root_y;
    const double bz = c->root_size*(double)c->N_root_z;
    uint64_t N = *Np;
    if (c->boundary==REBCU_BOUNDARY_OPEN){
        /* Order-preserving removal (particle.c:360-369); N_active shrinks for removed actives. */
        uint64_t w = 0;
        uint64_t Na = c->N_active;
        uint64_t removed_active = 0;
        for (uint64_t i=0;i<N;i++){
            const int out = p[i].x>bx/2. || p[i].x<-bx/2. || p[i].y>by/2. || p[i].y<-by/2.
                         || p[i].z>bz/2. || p[i].z<-bz/2.;
            if (out){
                if (Na!=REBCU_SIZE_MAX && i<Na) removed_active++;
            }else{
                if (w!=i) p[w] = p[i];
                w++;
            }
        }
        /* particle.c:336-343: removing the last remaining particle sets N=0 and does not touch N_active. */
        if (Na!=REBCU_SIZE_MAX){
            if (w==0 && N>0 && N-1<Na && removed_active>0) removed_active--; /* last removal skips the decrement */
            c->N_active = Na - removed_active;
        }
        *Np = w;
    }else if (c->boundary==REBCU_BOUNDARY_SHEAR){
        const double OMEGA = c->OMEGA;
        const double offp1 = -fmod(-1.5*OMEGA*bx*c->t + by/2., by) - by/2.;
        const double offm1 = -fmod( 1.5*OMEGA*bx*c->t - by/2., by) + by/2.;
        for (uint64_t i=0;i<N;i++){
            while (p[i].x> bx/2.){ p[i].x -= bx; p[i].y += offp1; p[i].vy += 3./2.*OMEGA*bx; }
            while (p[i].x<-bx/2.){ p[i].x += bx; p[i].y += offm1; p[i].vy -= 3./2.*OMEGA*bx; }
            while (p[i].y> by/2.) p[i].y -= by;
            while (p[i].y<-by/2.) p[i].y += by;
            while (p[i].z> bz/2.) p[i].z -= bz;
            while (p[i].z<-bz/2.) p[i].z += bz;
        }
    }else if (c->boundary==REBCU_BOUNDARY_PERIODIC){
        for (uint64_t i=0;i<N;i++){
            while (p[i].x> bx/2.) p[i].x -= bx;
            while (p[i].x<-bx/2.) p[i].x += bx;
            while (p[i].y> by/2.) p[i].y -= by;
            while (p[i].y<-by/2.) p[i].y += by;
            while (p[i].z> bz/2.) p[i].z -= bz;
            while (p[i].z<-bz/2.) p[i].z += bz;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Octree: src/tree.c:64-142 (insert), :147-207 (mass / centre of mass)                        */
/* Cells live in a growable array; children are array indices (0 = none, cell 0 is unused).    */
/* ------------------------------------------------------------------------------------------ */
typedef struct ocell {
    double x,y,z,w,m,mx,my,mz;
    double q[6];        /* mxx mxy mxz myy myz mzz: only with -DQUADRUPOLE (tree.h:43-50) */
    int pt;
    int rootbox;
    int kid[8];
} ocell;

typedef struct otree {
    ocell* c; size_t n, cap;
    int* root; int n_root;
} otree;

static int new_cell(otree* t){
    if (t->n==t->cap){ t->cap = t->cap? t->cap*2 : 1024; t->c = realloc(t->c, t->cap*sizeof(ocell)); }
    memset(&t->c[t->n], 0, sizeof(ocell));
    return (int)t->n++;
}

static int octant_of(const rebcu_particle* q, const ocell* cell){   /* tree.c:136-142 */
    int o = 0;
    if (q->x < cell->x) o += 1;
    if (q->y < cell->y) o += 2;
    if (q->z < cell->z) o += 4;
    return o;
}

static int rootbox_of(const rebcu_config* c, const rebcu_particle* q){   /* particle.c:119-126 */
    const double rs = c->root_size;
    int i = ((int)floor((q->x + rs*(double)c->N_root_x/2.)/rs)+c->N_root_x)%c->N_root_x;
    int j = ((int)floor((q->y + rs*(double)c->N_root_y/2.)/rs)+c->N_root_y)%c->N_root_y;
    int k = ((int)floor((q->z + rs*(double)c->N_root_z/2.)/rs)+c->N_root_z)%c->N_root_z;
    return (k*c->N_root_y+j)*c->N_root_x+i;
}

/* Creates the child cell of `parent` in octant o (tree.c:98-103), or a root cell (tree.c:86-97). */
static int make_cell(otree* t, const rebcu_config* c, const rebcu_particle* q, int pt, int parent, int o, int rootbox){
    int id = new_cell(t);
    ocell* n = &t->c[id];
    if (parent==0){
        const double rs = c->root_size;
        const double bx = rs*(double)c->N_root_x, by = rs*(double)c->N_root_y, bz = rs*(double)c->N_root_z;
        n->w = rs;
        int i = ((int)floor((q->x + bx/2.)/rs))%c->N_root_x;
        int j = ((int)floor((q->y + by/2.)/rs))%c->N_root_y;
        int k = ((int)floor((q->z + bz/2.)/rs))%c->N_root_z;
        n->x = -bx/2.+rs*(0.5+(double)i);
        n->y = -by/2.+rs*(0.5+(double)j);
        n->z = -bz/2.+rs*(0.5+(double)k);
    }else{
        const ocell* pc = &t->c[parent];
        n->w = pc->w/2.;
        n->x = pc->x + n->w/2.*((o>>0)%2==0?1.:-1);
        n->y = pc->y + n->w/2.*((o>>1)%2==0?1.:-1);
        n->z = pc->z + n->w/2.*((o>>2)%2==0?1.:-1);
    }
    n->pt = pt;
    n->rootbox = rootbox;
    return id;
}

static int tree_build(otree* t, const rebcu_config* c, const rebcu_particle* p, uint64_t N){
    memset(t, 0, sizeof(*t));
    if (c->root_size<=0.0)
        return orc_fail(REBCU_ERR_ROOT_SIZE, "Set root_size to a finite value to use a tree based gravity or collision solver.");
    t->n_root = c->N_root_x*c->N_root_y*c->N_root_z;
    t->root = calloc(t->n_root, sizeof(int));
    new_cell(t); /* index 0 = null */
    for (uint64_t i=0;i<N;i++){
        const rebcu_particle* q = &p[i];
        if (fabs(q->x)>c->root_size*(double)c->N_root_x/2. || fabs(q->y)>c->root_size*(double)c->N_root_y/2.
                || fabs(q->z)>c->root_size*(double)c->N_root_z/2.)
            return orc_fail(REBCU_ERR_OUTSIDE_BOX, "Particle is outside of simulation box. Cannot add to tree.");
        if (!isfinite(q->x) || !isfinite(q->y) || !isfinite(q->z))
            return orc_fail(REBCU_ERR_NONFINITE, "Particle has non-finite coordinates. Cannot add to tree.");
        const int rb = rootbox_of(c, q);
        if (t->root[rb]==0){ t->root[rb] = make_cell(t, c, q, (int)i, 0, 0, rb); continue; }
        int cur = t->root[rb];
        for(;;){
            if (t->c[cur].pt >= 0){
                /* Leaf: push the resident particle one level down, then continue with the new one. */
                const int old = t->c[cur].pt;
                const int o1 = octant_of(&p[old], &t->c[cur]);
                const int o2 = octant_of(q, &t->c[cur]);
                if (o1==o2 && q->x==p[old].x && q->y==p[old].y && q->z==p[old].z)
                    return orc_fail(REBCU_ERR_SAME_COORDINATES, "Cannot add two particles with the same coordinates to the tree.");
                int k1 = make_cell(t, c, &p[old], old, cur, o1, rb);
                if (t->c[k1].w<=0.0) return orc_fail(REBCU_ERR_CELL_SIZE_ZERO, "Tree cell has size zero.");
                t->c[cur].kid[o1] = k1;
                t->c[cur].pt = -2;
                if (o1==o2){ cur = k1; continue; }
                const int k2 = make_cell(t, c, q, (int)i, cur, o2, rb);
                t->c[cur].kid[o2] = k2;
                break;
            }else{
                t->c[cur].pt--;
                const int o = octant_of(q, &t->c[cur]);
                if (t->c[cur].kid[o]==0){
                    const int k = make_cell(t, c, q, (int)i, cur, o, rb);
                    if (t->c[k].w<=0.0) return orc_fail(REBCU_ERR_CELL_SIZE_ZERO, "Tree cell has size zero.");
                    t->c[cur].kid[o] = k;
                    break;
                }
                cur = t->c[cur].kid[o];
            }
        }
    }
    return 0;
}

/* Post-order mass / centre of mass, children combined in octant order (tree.c:156-206). */
static void tree_moments(otree* t, const rebcu_particle* p, int id, int quadrupole){
    ocell* n = &t->c[id];
    for (int k=0;k<6;k++) n->q[k] = 0.;                      /* tree.c:148-155 */
    if (n->pt < 0){
        double m=0., mx=0., my=0., mz=0.;
        for (int o=0;o<8;o++){
            const int k = t->c[id].kid[o];
            if (!k) continue;
            tree_moments(t, p, k, quadrupole);
            const ocell* d = &t->c[k];
            const double dm = d->m;
            mx += d->mx*dm; my += d->my*dm; mz += d->mz*dm; m += dm;
        }
        n = &t->c[id];
        if (m>0){ mx /= m; my /= m; mz /= m; }
        n->m = m; n->mx = mx; n->my = my; n->mz = mz;
        if (quadrupole){                                      /* tree.c:180-197, Hernquist 1987 */
            for (int o=0;o<8;o++){
                const int k = n->kid[o];
                if (!k) continue;
                const ocell* d = &t->c[k];
                const double d_m = d->m;
                const double qx = d->mx - n->mx, qy = d->my - n->my, qz = d->mz - n->mz;
                const double qr2 = qx*qx + qy*qy + qz*qz;
                n->q[0] += d->q[0] + d_m*(3.*qx*qx - qr2);
                n->q[1] += d->q[1] + d_m*3.*qx*qy;
                n->q[2] += d->q[2] + d_m*3.*qx*qz;
                n->q[3] += d->q[3] + d_m*(3.*qy*qy - qr2);
                n->q[4] += d->q[4] + d_m*3.*qy*qz;
            }
            n->q[5] = -n->q[0] -n->q[3];
        }
    }else{
        n->m = p[n->pt].m; n->mx = p[n->pt].x; n->my = p[n->pt].y; n->mz = p[n->pt].z;
    }
}

/* Flatten to depth-first pre-order with skip links. */
static size_t tree_flatten(const otree* t, int id, int depth, rebcu_treecell* out, size_t cap, size_t idx, double* quad){
    const size_t me = idx++;
    const ocell* n = &t->c[id];
    for (int o=0;o<8;o++) if (n->kid[o]) idx = tree_flatten(t, n->kid[o], depth+1, out, cap, idx, quad);
    if (me<cap && quad) for (int k=0;k<6;k++) quad[6*me+k] = n->q[k];
    if (me<cap){
        rebcu_treecell* q = &out[me];
        q->x=n->x; q->y=n->y; q->z=n->z; q->w=n->w; q->m=n->m; q->mx=n->mx; q->my=n->my; q->mz=n->mz;
        q->pt=n->pt; q->skip=(int32_t)idx; q->depth=depth; q->rootbox=n->rootbox;
    }
    return idx;
}

static void tree_free(otree* t){ free(t->c); free(t->root); memset(t,0,sizeof(*t)); }

/* Builds the flattened tree (with moments) for the given particles. */
static int build_flat_q(const rebcu_config* c, const rebcu_particle* p, uint64_t N,
                        rebcu_treecell** cells, size_t* n_cells, double** quad_out){
    otree t;
    int err = tree_build(&t, c, p, N);
    if (quad_out) *quad_out = NULL;
    if (err){ tree_free(&t); *cells = NULL; *n_cells = 0; return err; }
    size_t total = t.n - 1;
    rebcu_treecell* out = malloc(sizeof(rebcu_treecell)*(total?total:1));
    double* quad = (quad_out && c->quadrupole) ? malloc(sizeof(double)*6*(total?total:1)) : NULL;
    size_t idx = 0;
    for (int rb=0; rb<t.n_root; rb++){
        if (!t.root[rb]) continue;
        tree_moments(&t, p, t.root[rb], c->quadrupole);
        idx = tree_flatten(&t, t.root[rb], 0, out, total, idx, quad);
    }
    tree_free(&t);
    *cells = out; *n_cells = idx;
    if (quad_out) *quad_out = quad;
    return 0;
}

static int build_flat(const rebcu_config* c, const rebcu_particle* p, uint64_t N,
                      rebcu_treecell** cells, size_t* n_cells){
    return build_flat_q(c, p, N, cells, n_cells, NULL);
}

int orc_tree_dump(rebcu_config* c, rebcu_particle* p, uint64_t N,
                  rebcu_treecell* out, uint64_t cap, uint64_t* n_cells){
    rebcu_treecell* cells; size_t n;
    int err = build_flat(c, p, N, &cells, &n);
    *n_cells = n;
    if (!err){
        memcpy(out, cells, sizeof(rebcu_treecell)*(n<cap?n:cap));
    }
    free(cells);
    return err;
}

/* Barnes-Hut walk over the pre-order array: gravity.c:84-99 and tree.c:275-328.
 * A cell is opened iff  w*w > opening_angle2 * r2  (tree.c:284); the distance uses the unsoftened r2. */
static void tree_walk_one(const rebcu_config* c, const rebcu_treecell* cells, size_t n_cells, const double* quad,
                          int self, double px, double py, double pz, double* a){
    const double G = c->G, soft2 = c->softening*c->softening, th2 = c->opening_angle2;
    size_t k = 0;
    while (k<n_cells){
        const rebcu_treecell* n = &cells[k];
        const double dx = px - n->mx, dy = py - n->my, dz = pz - n->mz;
        const double r2 = dx*dx + dy*dy + dz*dz;
        if (n->pt < 0){
            if (n->w*n->w > th2*r2){ k++; continue; }
        }else if (n->pt == self){ k = n->skip; continue; }
        const double rr = sqrt(r2 + soft2);
        const double pre = -G/(rr*rr*rr)*n->m;
        if (quad && n->pt < 0){                               /* tree.c:293-303 */
            const double* q = quad + 6*k;                     /* mxx mxy mxz myy myz mzz */
            double qpre = G/(rr*rr*rr*rr*rr);
            a[0] += qpre*(dx*q[0] + dy*q[1] + dz*q[2]);
            a[1] += qpre*(dx*q[1] + dy*q[3] + dz*q[4]);
            a[2] += qpre*(dx*q[2] + dy*q[4] + dz*q[5]);
            double mrr = dx*dx*q[0] + dy*dy*q[3] + dz*dz*q[5]
                       + 2.*dx*dy*q[1] + 2.*dx*dz*q[2] + 2.*dy*dz*q[4];
            qpre *= -5.0/(2.0*rr*rr)*mrr;
            a[0] += (qpre + pre) * dx;
            a[1] += (qpre + pre) * dy;
            a[2] += (qpre + pre) * dz;
        }else{
            a[0] += pre*dx; a[1] += pre*dy; a[2] += pre*dz;
        }
        k = n->skip;
    }
}

static int gravity_tree(rebcu_config* c, rebcu_particle* p, uint64_t* Np){
    boundary_check(c, p, Np);                                   /* gravity.c:56 */
    const uint64_t N = *Np;
    rebcu_treecell* cells; size_t n_cells; double* quad = NULL;
    int err = build_flat_q(c, p, N, &cells, &n_cells, &quad);
    if (err) return err;
    const int ngb = (2*c->N_ghost_x+1)*(2*c->N_ghost_y+1)*(2*c->N_ghost_z+1);
    rebcu_vec6d* gbs = malloc(sizeof(rebcu_vec6d)*ngb);
    int n=0;
    for (int gx=-c->N_ghost_x; gx<=c->N_ghost_x; gx++)
        for (int gy=-c->N_ghost_y; gy<=c->N_ghost_y; gy++)
            for (int gz=-c->N_ghost_z; gz<=c->N_ghost_z; gz++)
                gbs[n++] = ghostbox(c, gx, gy, gz);
#pragma omp parallel for schedule(dynamic,64)
    for (uint64_t i=0;i<N;i++){
        double a[3] = {0.,0.,0.};
        for (int g=0; g<ngb; g++)
            tree_walk_one(c, cells, n_cells, quad, (int)i, gbs[g].x+p[i].x, gbs[g].y+p[i].y, gbs[g].z+p[i].z, a);
        p[i].ax=a[0]; p[i].ay=a[1]; p[i].az=a[2];
    }
    free(gbs); free(cells); free(quad);
    return 0;
}

/* reb_simulation_update_acceleration, simulation.c:640-689 */
static int update_acceleration(rebcu_config* c, rebcu_particle* p, uint64_t* Np){
    switch (c->gravity){
        case REBCU_GRAVITY_NONE:
            for (uint64_t i=0;i<*Np;i++){ p[i].ax=0; p[i].ay=0; p[i].az=0; }
            return 0;
        case REBCU_GRAVITY_BASIC: gravity_basic(c, p, *Np); return 0;
        case REBCU_GRAVITY_COMPENSATED: gravity_compensated(c, p, *Np, NULL); return 0;
        case REBCU_GRAVITY_TREE: return gravity_tree(c, p, Np);
        default: return orc_fail(REBCU_ERR_ARG, "Gravity calculation not yet implemented.");
    }
}

int orc_gravity(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    orc_errbuf[0]=0;
    return update_acceleration(c, p, N);
}

/* Force evaluation that also returns r->gravity_cs (3 doubles per particle); COMPENSATED only. */
int orc_gravity_cs(rebcu_config* c, rebcu_particle* p, uint64_t* N, double* cs_out){
    orc_errbuf[0]=0;
    if (c->gravity!=REBCU_GRAVITY_COMPENSATED){ snprintf(orc_errbuf, sizeof(orc_errbuf), "gravity_cs needs REB_GRAVITY_COMPENSATED"); return -2; }
    gravity_compensated(c, p, *N, cs_out);
    return 0;
}

int orc_gravity_timed(rebcu_config* c, rebcu_particle* p, uint64_t* N, int n_evals, double* sec){
    struct timespec t0, t1; int err=0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k=0;k<n_evals && !err;k++) err = update_acceleration(c, p, N);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *sec = ((t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec))/n_evals;
    return err;
}

int orc_boundary_check(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    orc_errbuf[0]=0;
    boundary_check(c, p, N);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Integrators: src/integrator_leapfrog.c:72-209, src/integrator_sei.c:86-174                  */
/* ------------------------------------------------------------------------------------------ */
static void lf_drift(rebcu_config* c, rebcu_particle* p, uint64_t N, double dt){
    for (uint64_t i=0;i<N;i++){ p[i].x += dt*p[i].vx; p[i].y += dt*p[i].vy; p[i].z += dt*p[i].vz; }
    c->t += dt;
}
static void lf_kick(rebcu_particle* p, uint64_t N, double dt){
    for (uint64_t i=0;i<N;i++){ p[i].vx += dt*p[i].ax; p[i].vy += dt*p[i].ay; p[i].vz += dt*p[i].az; }
}

/* Composition coefficients (integrator_leapfrog.c:68-70). */
static const double LF4 = 0.675603595979828817023843904485;
static const double LF6[5] = {0.1867, 0.5554970237124784, 0.1294669489134754, -0.843265623387734, 0.9432033015235604};
static const double LF8[9] = {0.128865979381443, 0.581514087105251, -0.410175371469850, 0.1851469357165877, -0.4095523434208514, 0.1444059410800120, 0.2783355003936797, 0.3149566839162949, -0.6269948254051343979};

/* The drift/kick schedule of one leapfrog step as two coefficient lists: n_kick kicks, n_kick+1 drifts.
 * The products are formed exactly as the reference writes them (integrator_leapfrog.c:102-203). */
static int lf_schedule(int order, double dt, double* drift, double* kick){
    int nk = 0;
    if (order==2){
        drift[0] = dt*0.5; kick[0] = dt; drift[1] = dt*0.5; nk = 1;
    }else if (order==4){
        drift[0] = dt*LF4;         kick[0] = dt*2.*LF4;
        drift[1] = dt*(0.5-LF4);   kick[1] = dt*(1.-4.*LF4);
        drift[2] = dt*(0.5-LF4);   kick[2] = dt*2.*LF4;
        drift[3] = dt*LF4; nk = 3;
    }else if (order==6 || order==8){
        const double* a = (order==6)?LF6:LF8;
        const int s = (order==6)?5:9;       /* palindromic: a0..a(s-1)..a0 */
        nk = 2*s-1;
        for (int k=0;k<nk;k++){
            const int idx = (k<s)?k:(2*s-2-k);
            kick[k] = dt*a[idx];
        }
        drift[0] = dt*a[0]*0.5;
        for (int k=1;k<nk;k++){
            const int lo = (k<s)?(k-1):(2*s-2-k);   /* smaller index of the adjacent pair */
            drift[k] = dt*(a[lo]+a[lo+1])*0.5;
        }
        drift[nk] = dt*a[0]*0.5;
    }else return -1;
    return nk;
}

static int leapfrog_step(rebcu_config* c, rebcu_particle* p, uint64_t* Np){
    c->gravity_ignore_terms = REBCU_IGNORE_TERMS_NONE;
    double drift[20], kick[20];
    const int order = c->leapfrog_order ? c->leapfrog_order : 2;
    const int nk = lf_schedule(order, c->dt, drift, kick);
    if (nk<0) return orc_fail(REBCU_ERR_LEAPFROG_ORDER, "Leapfrog order not supported.");
    for (int k=0;k<nk;k++){
        lf_drift(c, p, *Np, drift[k]);
        int err = update_acceleration(c, p, Np);
        if (err) return err;
        lf_kick(p, *Np, kick[k]);
    }
    lf_drift(c, p, *Np, drift[nk]);
    c->dt_last_done = c->dt;
    return 0;
}

typedef struct { double sindt, tandt, sindtz, tandtz; } sei_consts;

static void sei_h012(double dt, const sei_consts* s, rebcu_particle* q, double OMEGA, double OMEGAZ){
    /* vertical: rotation as three shears (integrator_sei.c:129-139) */
    const double zx = q->z*OMEGAZ;
    const double zy = q->vz;
    const double zt1 = zx - s->tandtz*zy;
    const double zyt = s->sindtz*zt1 + zy;
    const double zxt = zt1 - s->tandtz*zyt;
    q->z = zxt/OMEGAZ;
    q->vz = zyt;
    /* horizontal epicycle (integrator_sei.c:141-156) */
    const double aO = 2.*q->vy + 4.*q->x*OMEGA;
    const double bO = q->y*OMEGA - 2.*q->vx;
    const double ys = (q->y*OMEGA-bO)/2.;
    const double xs = (q->x*OMEGA-aO);
    const double xst1 = xs - s->tandt*ys;
    const double yst  = s->sindt*xst1 + ys;
    const double xst  = xst1 - s->tandt*yst;
    q->x  = (xst+aO)/OMEGA;
    q->y  = (yst*2.+bO)/OMEGA - 3./4.*aO*dt;
    q->vx = yst;
    q->vy = -xst*2. - 3./2.*aO;
}

static int sei_step(rebcu_config* c, rebcu_particle* p, uint64_t* Np){
    c->gravity_ignore_terms = REBCU_IGNORE_TERMS_NONE;
    if (c->OMEGAZ==-1) c->OMEGAZ = c->OMEGA;                 /* integrator_sei.c:93-95 */
    sei_consts s;
    s.sindt  = sin(c->OMEGA*(-c->dt/2.));
    s.tandt  = tan(c->OMEGA*(-c->dt/4.));
    s.sindtz = sin(c->OMEGAZ*(-c->dt/2.));
    s.tandtz = tan(c->OMEGAZ*(-c->dt/4.));
    for (uint64_t i=0;i<*Np;i++) sei_h012(c->dt, &s, &p[i], c->OMEGA, c->OMEGAZ);
    c->t += c->dt/2.;
    int err = update_acceleration(c, p, Np);
    if (err) return err;
    for (uint64_t i=0;i<*Np;i++){
        p[i].vx += p[i].ax*c->dt; p[i].vy += p[i].ay*c->dt; p[i].vz += p[i].az*c->dt;   /* phi1, :168-174 */
        sei_h012(c->dt, &s, &p[i], c->OMEGA, c->OMEGAZ);
    }
    c->t += c->dt/2.;
    c->dt_last_done = c->dt;
    return 0;
}

static int integrator_step(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    switch (c->integrator){
        case REBCU_INTEGRATOR_LEAPFROG: return leapfrog_step(c, p, N);
        case REBCU_INTEGRATOR_SEI: return sei_step(c, p, N);
        default: c->t += c->dt; c->dt_last_done = c->dt; return 0;   /* reb_integrator_none, rebound.c */
    }
}

int orc_integrator_step(rebcu_config* c, rebcu_particle* p, uint64_t* N){
    orc_errbuf[0]=0;
    return integrator_step(c, p, N);
}

/* ------------------------------------------------------------------------------------------ */
/* Collision search: src/collision.c:49-331, :422-503                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct clist { rebcu_collision* v; size_t n, cap; } clist;
static void clist_push(clist* l, uint64_t p1, uint64_t p2, rebcu_vec6d gb, uint64_t ri){
    if (l->n==l->cap){ l->cap = l->cap? l->cap*2:32; l->v = realloc(l->v, l->cap*sizeof(rebcu_collision)); }
    rebcu_collision* q = &l->v[l->n++];
    memset(q, 0, sizeof(*q));
    q->p1=p1; q->p2=p2; q->gb=gb; q->ri=ri;
}

static int overlapping_and_approaching(const rebcu_vec6d* s, double r1, const rebcu_particle* q){
    const double dx = s->x - q->x, dy = s->y - q->y, dz = s->z - q->z;
    const double sr = r1 + q->r;
    const double r2 = dx*dx + dy*dy + dz*dz;
    if (r2 > sr*sr) return 0;
    const double dvx = s->vx - q->vx, dvy = s->vy - q->vy, dvz = s->vz - q->vz;
    if (dvx*dx + dvy*dy + dvz*dz > 0) return 0;
    return 1;
}

/* Straight-line trajectories over the last step come closer than the sum of the radii
 * (collision.c:155-177 for LINE, :513-535 for LINETREE; same arithmetic). */
static int trajectories_overlap(const rebcu_vec6d* s, double r1, const rebcu_particle* q, double dt_last_done){
    const double dx1 = s->x - q->x, dy1 = s->y - q->y, dz1 = s->z - q->z;
    const double r1sq = (dx1*dx1 + dy1*dy1 + dz1*dz1);
    const double dvx1 = s->vx - q->vx, dvy1 = s->vy - q->vy, dvz1 = s->vz - q->vz;
    const double dx2 = dx1 - dt_last_done*dvx1, dy2 = dy1 - dt_last_done*dvy1, dz2 = dz1 - dt_last_done*dvz1;
    const double r2sq = (dx2*dx2 + dy2*dy2 + dz2*dz2);
    const double t_closest = (dx1*dvx1 + dy1*dvy1 + dz1*dvz1)/(dvx1*dvx1 + dvy1*dvy1 + dvz1*dvz1);
    double rmin2 = (r1sq > r2sq) ? r2sq : r1sq;                      /* MIN(r1,r2), collision.c:43 */
    if (t_closest/dt_last_done>=0. && t_closest/dt_last_done<=1.){
        const double dx3 = dx1 - t_closest*dvx1, dy3 = dy1 - t_closest*dvy1, dz3 = dz1 - t_closest*dvz1;
        const double r3sq = (dx3*dx3 + dy3*dy3 + dz3*dz3);
        rmin2 = (rmin2 > r3sq) ? r3sq : rmin2;
    }
    const double rsum = r1 + q->r;
    return !(rmin2 > rsum*rsum);
}

/* Radius of the second largest particle, first-index-wins on ties (simulation.c:782-798). */
static double second_largest_radius(const rebcu_particle* p, uint64_t N){
    double l1=-1.0, l2=-1.0; int have2 = 0; int have1 = 0;
    for (uint64_t i=0;i<N;i++){
        if (p[i].r > l1){ l2 = l1; have2 = have1; l1 = p[i].r; have1 = 1; }
        else if (p[i].r > l2){ l2 = p[i].r; have2 = 1; }
    }
    return have2 ? l2 : 0.;
}

/* map / N_map / N_targets: r->map, r->N_map, r->N_targets (collision.c:53-58).  map==NULL: all N particles are
 * projectiles; N_targets==REBCU_SIZE_MAX: as many targets as projectiles.  DIRECT maps projectile and target
 * slots (:78,93) and skips equal SLOTS (:92); LINE maps both sides but ignores N_targets (:140,153); the tree
 * modes only cut the projectile loop to the first N_projectiles particles, unmapped (:229,:286). */
static int collision_search(const rebcu_config* c, const rebcu_particle* p, uint64_t N,
                            const uint64_t* map, uint64_t N_map, uint64_t N_targets_in, clist* out){
    const uint64_t NP = map ? N_map : N;
    const uint64_t NT = N_targets_in!=REBCU_SIZE_MAX ? N_targets_in : NP;
    const int gx1 = c->N_ghost_x>1?1:c->N_ghost_x;
    const int gy1 = c->N_ghost_y>1?1:c->N_ghost_y;
    const int gz1 = c->N_ghost_z>1?1:c->N_ghost_z;
    if (c->collision==REBCU_COLLISION_DIRECT){
        /* ghost box outermost, then projectile, then target (collision.c:70-121) */
        for (int gx=-gx1; gx<=gx1; gx++) for (int gy=-gy1; gy<=gy1; gy++) for (int gz=-gz1; gz<=gz1; gz++){
            const rebcu_vec6d gb = ghostbox(c, gx, gy, gz);
            for (uint64_t i=0;i<NP;i++){
                const uint64_t ip = map ? map[i] : i;
                rebcu_vec6d s = gb;
                s.x += p[ip].x; s.y += p[ip].y; s.z += p[ip].z; s.vx += p[ip].vx; s.vy += p[ip].vy; s.vz += p[ip].vz;
                for (uint64_t j=0;j<NT;j++){
                    if (i==j) continue;
                    const uint64_t jp = map ? map[j] : j;
                    if (overlapping_and_approaching(&s, p[ip].r, &p[jp])) clist_push(out, ip, jp, gb, 0);
                }
            }
        }
        return 0;
    }
    if (c->collision==REBCU_COLLISION_TREE){
        /* projectile outermost, then ghost box, root box, depth-first descent (collision.c:229-266) */
        rebcu_treecell* cells; size_t n_cells;
        int err = build_flat(c, p, N, &cells, &n_cells);
        if (err) return err;
        const double r2nd = second_largest_radius(p, N);
        for (uint64_t i=0;i<NP;i++){
            for (int gx=-gx1; gx<=gx1; gx++) for (int gy=-gy1; gy<=gy1; gy++) for (int gz=-gz1; gz<=gz1; gz++){
                const rebcu_vec6d gb = ghostbox(c, gx, gy, gz);
                rebcu_vec6d s = gb;
                s.x += p[i].x; s.y += p[i].y; s.z += p[i].z; s.vx += p[i].vx; s.vy += p[i].vy; s.vz += p[i].vz;
                size_t k = 0;
                while (k<n_cells){
                    const rebcu_treecell* n = &cells[k];
                    if (n->pt>=0){
                        if ((uint64_t)n->pt!=i && overlapping_and_approaching(&s, p[i].r, &p[n->pt]))
                            clist_push(out, i, (uint64_t)n->pt, gb, (uint64_t)n->rootbox);
                        k = n->skip;
                    }else{
                        const double dx = s.x - n->x, dy = s.y - n->y, dz = s.z - n->z;
                        const double r2 = dx*dx + dy*dy + dz*dz;
                        const double rp = p[i].r + r2nd + 0.86602540378443*n->w;   /* collision.c:492 */
                        k = (r2 < rp*rp) ? k+1 : (size_t)n->skip;
                    }
                }
            }
        }
        free(cells);
        return 0;
    }
    if (c->collision==REBCU_COLLISION_LINE){
        /* ghost box outermost, then i, then j > i (collision.c:132-194) */
        for (int gx=-gx1; gx<=gx1; gx++) for (int gy=-gy1; gy<=gy1; gy++) for (int gz=-gz1; gz<=gz1; gz++){
            const rebcu_vec6d gb = ghostbox(c, gx, gy, gz);
            for (uint64_t i=0;i<NP;i++){
                const uint64_t ip = map ? map[i] : i;
                rebcu_vec6d s = gb;
                s.x += p[ip].x; s.y += p[ip].y; s.z += p[ip].z; s.vx += p[ip].vx; s.vy += p[ip].vy; s.vz += p[ip].vz;
                for (uint64_t j=i+1;j<NP;j++){
                    const uint64_t jp = map ? map[j] : j;
                    if (trajectories_overlap(&s, p[ip].r, &p[jp], c->dt_last_done)) clist_push(out, ip, jp, gb, 0);
                }
            }
        }
        return 0;
    }
    if (c->collision==REBCU_COLLISION_LINETREE){
        /* collision.c:270-331 with the descent of :506-568 */
        double vmax2 = 0.;
        for (uint64_t i=0;i<NP;i++){
            const double v2 = p[i].vx*p[i].vx + p[i].vy*p[i].vy + p[i].vz*p[i].vz;
            vmax2 = (vmax2 > v2) ? vmax2 : v2;                         /* MAX(vmax2, v2), collision.c:44 */
        }
        const double maxdrift = c->dt_last_done*sqrt(vmax2);
        rebcu_treecell* cells; size_t n_cells;
        int err = build_flat(c, p, N, &cells, &n_cells);
        if (err) return err;
        for (uint64_t i=0;i<NP;i++){
            const double reach = p[i].r + c->dt_last_done*sqrt(p[i].vx*p[i].vx + p[i].vy*p[i].vy + p[i].vz*p[i].vz);
            for (int gx=-gx1; gx<=gx1; gx++) for (int gy=-gy1; gy<=gy1; gy++) for (int gz=-gz1; gz<=gz1; gz++){
                const rebcu_vec6d gb = ghostbox(c, gx, gy, gz);
                rebcu_vec6d s = gb;
                s.x += p[i].x; s.y += p[i].y; s.z += p[i].z; s.vx += p[i].vx; s.vy += p[i].vy; s.vz += p[i].vz;
                size_t k = 0;
                while (k<n_cells){
                    const rebcu_treecell* n = &cells[k];
                    if (n->pt>=0){
                        if ((uint64_t)n->pt!=i && trajectories_overlap(&s, p[i].r, &p[n->pt], c->dt_last_done))
                            clist_push(out, i, (uint64_t)n->pt, gb, (uint64_t)n->rootbox);
                        k = n->skip;
                    }else{
                        const double dx = s.x - n->x, dy = s.y - n->y, dz = s.z - n->z;
                        const double r2 = dx*dx + dy*dy + dz*dz;
                        const double rp = reach + maxdrift + 0.86602540378443*n->w;   /* collision.c:557 */
                        k = (r2 < rp*rp) ? k+1 : (size_t)n->skip;
                    }
                }
            }
        }
        free(cells);
        return 0;
    }
    return 0;
}

int orc_collision_search(rebcu_config* c, rebcu_particle* p, uint64_t N,
                         rebcu_collision* out, uint64_t cap, uint64_t* n_found){
    orc_errbuf[0]=0;
    clist l = {0,0,0};
    int err = collision_search(c, p, N, NULL, 0, REBCU_SIZE_MAX, &l);
    *n_found = l.n;
    if (out && l.n) memcpy(out, l.v, sizeof(rebcu_collision)*(l.n<cap?l.n:cap));
    free(l.v);
    return err;
}

/* The same search restricted by r->map / r->N_map / r->N_targets (what MERCURIUS and TRACE set around their
 * encounter steps, integrator_mercurius.c:429, integrator_trace.c:820). */
int orc_collision_search_subset(rebcu_config* c, rebcu_particle* p, uint64_t N,
                                const uint64_t* map, uint64_t N_map, uint64_t N_targets,
                                rebcu_collision* out, uint64_t cap, uint64_t* n_found){
    orc_errbuf[0]=0;
    clist l = {0,0,0};
    int err = collision_search(c, p, N, map, N_map, N_targets, &l);
    *n_found = l.n;
    if (out && l.n) memcpy(out, l.v, sizeof(rebcu_collision)*(l.n<cap?l.n:cap));
    free(l.v);
    return err;
}

/* ------------------------------------------------------------------------------------------ */
/* Collision resolve (host side in the product as well): collision.c:336-404, :573-665          */
/* ------------------------------------------------------------------------------------------ */
typedef struct resolve_ctx {
    int kind;                          /* 0 none, 1 hard sphere eps=1, 2 hard sphere + Bridges law */
    double minimum_collision_velocity;
    double plog; long long log_n;
} resolve_ctx;

static double restitution(const resolve_ctx* rc, double v){
    if (rc->kind!=2) return 1.;
    double eps = 0.32*pow(fabs(v)*100.,-0.234);     /* examples/shearing_sheet/problem.c:96-103 */
    if (eps>1) eps=1;
    if (eps<0) eps=0;
    return eps;
}

static void resolve_hardsphere(resolve_ctx* rc, rebcu_particle* P, const rebcu_collision* col){
    const rebcu_particle p1 = P[col->p1], p2 = P[col->p2];
    const rebcu_vec6d gb = col->gb;
    const double x21 = p1.x + gb.x - p2.x, y21 = p1.y + gb.y - p2.y, z21 = p1.z + gb.z - p2.z;
    const double rp = p1.r + p2.r;
    const double oldvyouter = (x21>0) ? p1.vy : p2.vy;
    if (rp*rp < x21*x21 + y21*y21 + z21*z21) return;
    const double vx21 = p1.vx + gb.vx - p2.vx, vy21 = p1.vy + gb.vy - p2.vy, vz21 = p1.vz + gb.vz - p2.vz;
    if (vx21*x21 + vy21*y21 + vz21*z21 > 0) return;
    const double theta = atan2(z21,y21), stheta = sin(theta), ctheta = cos(theta);
    const double vy21n = ctheta*vy21 + stheta*vz21;
    const double y21n = ctheta*y21 + stheta*z21;
    const double phi = atan2(y21n,x21), cphi = cos(phi), sphi = sin(phi);
    const double vx21nn = cphi*vx21 + sphi*vy21n;
    const double eps = restitution(rc, vx21nn);
    double dvx2 = -(1.0+eps)*vx21nn;
    const double minr = (p1.r>p2.r)?p2.r:p1.r;
    const double maxr = (p1.r<p2.r)?p2.r:p1.r;
    double mindv = minr*rc->minimum_collision_velocity;
    const double rr = sqrt(x21*x21 + y21*y21 + z21*z21);
    mindv *= 1.-(rr - maxr)/minr;
    if (mindv>maxr*rc->minimum_collision_velocity) mindv = maxr*rc->minimum_collision_velocity;
    if (dvx2<mindv) dvx2 = mindv;
    const double dvx2n = cphi*dvx2, dvy2n = sphi*dvx2;
    const double dvy2nn = ctheta*dvy2n, dvz2nn = stheta*dvy2n;
    const double p2pf = p1.m/(p1.m+p2.m);
    P[col->p2].vx -= p2pf*dvx2n; P[col->p2].vy -= p2pf*dvy2nn; P[col->p2].vz -= p2pf*dvz2nn;
    const double p1pf = p2.m/(p1.m+p2.m);
    P[col->p1].vx += p1pf*dvx2n; P[col->p1].vy += p1pf*dvy2nn; P[col->p1].vz += p1pf*dvz2nn;
    if (x21>0) rc->plog += -fabs(x21)*(oldvyouter-P[col->p1].vy)*p1.m;
    else       rc->plog += -fabs(x21)*(oldvyouter-P[col->p2].vy)*p2.m;
    rc->log_n++;
}

/* reb_simulation_steps without callbacks: simulation.c:504-603.  The shuffle uses glibc rand_r on a
 * seed that starts at 42 for every call (the harness does the same with r->rand_seed). */
int orc_steps(rebcu_config* c, rebcu_particle* p, uint64_t* N, uint64_t n_steps,
              int resolve, double minimum_collision_velocity, double* aux){
    orc_errbuf[0]=0;
    resolve_ctx rc = {resolve, minimum_collision_velocity, 0., 0};
    unsigned int seed = 42;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int err = 0;
    for (uint64_t s=0; s<n_steps && !err; s++){
        err = integrator_step(c, p, N);
        if (err) break;
        boundary_check(c, p, N);
        if (c->collision!=REBCU_COLLISION_NONE){
            clist l = {0,0,0};
            err = collision_search(c, p, *N, NULL, 0, REBCU_SIZE_MAX, &l);
            for (size_t i=0;i<l.n;i++){             /* collision.c:337-342 */
                size_t j = rand_r(&seed)%l.n;
                rebcu_collision t = l.v[i]; l.v[i] = l.v[j]; l.v[j] = t;
            }
            if (resolve) for (size_t i=0;i<l.n;i++) resolve_hardsphere(&rc, p, &l.v[i]);
            free(l.v);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (aux){ aux[0] = (double)rc.log_n; aux[1] = rc.plog; aux[2] = (t1.tv_sec-t0.tv_sec) + 1e-9*(t1.tv_nsec-t0.tv_nsec); }
    return err;
}

/* reb_simulation_energy, tools.c:108-162 */
double orc_energy(rebcu_config* c, rebcu_particle* p, uint64_t N){
    const uint64_t Na = (c->N_active==REBCU_SIZE_MAX)?N:c->N_active;
    const uint64_t Ni = (c->testparticle_type==0)?Na:N;
    double ek=0., ep=0.;
    for (uint64_t i=0;i<Ni;i++) ek += 0.5*p[i].m*(p[i].vx*p[i].vx + p[i].vy*p[i].vy + p[i].vz*p[i].vz);
    for (uint64_t i=0;i<Na;i++) for (uint64_t j=i+1;j<Ni;j++){
        const double dx = p[i].x-p[j].x, dy = p[i].y-p[j].y, dz = p[i].z-p[j].z;
        ep -= c->G*p[j].m*p[i].m/sqrt(dx*dx + dy*dy + dz*dz);
    }
    return ek + ep;
}

/* reb_simulation_com / reb_particle_com_of_pair, tools.c:376-408: out = {m,x,y,z,vx,vy,vz,ax,ay,az} */
void orc_com(rebcu_config* c, rebcu_particle* p, uint64_t N, double* out){
    (void)c;
    double m=0., q[9]={0.};
    for (uint64_t i=0;i<N;i++){
        const double v[9] = {p[i].x,p[i].y,p[i].z,p[i].vx,p[i].vy,p[i].vz,p[i].ax,p[i].ay,p[i].az};
        for (int k=0;k<9;k++) q[k] = q[k]*m + v[k]*p[i].m;
        m += p[i].m;
        if (m>0.) for (int k=0;k<9;k++) q[k] /= m;
    }
    out[0] = m;
    for (int k=0;k<9;k++) out[1+k] = q[k];
}

/* reb_simulation_angular_momentum, tools.c:164-174 */
void orc_angular_momentum(rebcu_config* c, rebcu_particle* p, uint64_t N, double* out){
    (void)c;
    double lx=0., ly=0., lz=0.;
    for (uint64_t i=0;i<N;i++){
        lx += p[i].m*(p[i].y*p[i].vz - p[i].z*p[i].vy);
        ly += p[i].m*(p[i].z*p[i].vx - p[i].x*p[i].vz);
        lz += p[i].m*(p[i].x*p[i].vy - p[i].y*p[i].vx);
    }
    out[0]=lx; out[1]=ly; out[2]=lz;
}

/* Exit conditions of run_heartbeat (src/simulation.c:242-272): returns the status the reference ends up with,
 * 4 = REB_STATUS_ESCAPE, 3 = REB_STATUS_ENCOUNTER (tested second, so it wins), 0 = neither. */
int orc_exit_check(rebcu_config* c, rebcu_particle* p, uint64_t N, double exit_max_distance, double exit_min_distance){
    (void)c;
    int status = 0;
    if (exit_max_distance){
        const double max2 = exit_max_distance*exit_max_distance;
        for (uint64_t i=0;i<N;i++){
            const double r2 = p[i].x*p[i].x + p[i].y*p[i].y + p[i].z*p[i].z;
            if (r2>max2) status = 4;
        }
    }
    if (exit_min_distance){
        const double min2 = exit_min_distance*exit_min_distance;
        for (uint64_t i=0;i<N;i++) for (uint64_t j=0;j<i;j++){
            const double x = p[i].x-p[j].x, y = p[i].y-p[j].y, z = p[i].z-p[j].z;
            const double r2 = x*x + y*y + z*z;
            if (r2<min2) status = 3;
        }
    }
    return status;
}

int orc_openmp_threads(void){
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n){
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
