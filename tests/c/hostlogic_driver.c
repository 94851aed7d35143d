/*
 * hostlogic_driver.c -- call sequences that stress the HOST side of the drop-in (particle arrays that grow, shrink and
 * move between calls, integrator switches, copies of resident simulations, several simulations in one process and in
 * several threads, error paths), written against the reference's public API only.  tests/test_hostlogic_cpu.py
 * runs it once on the unmodified reference and once on the drop-in with the mock engine (all residency modes) and
 * compares the dumps bit for bit.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include "rebound.h"

static FILE* out;
static void dump(const char* tag, struct reb_simulation* r){
    double hdr[5] = {(double)r->N, r->t, r->dt_last_done, (double)r->status, (double)r->steps_done};
    fwrite(tag, 1, 4, out);
    fwrite(hdr, sizeof(double), 5, out);
    for (size_t i=0;i<r->N;i++) fwrite(&r->particles[i], sizeof(double), 11, out);
}

static void cloud(struct reb_simulation* r, int n, double vel, double radius){
    for (int i=0;i<n;i++){
        struct reb_particle p = {0};
        p.x = reb_random_uniform(r,-1.,1.); p.y = reb_random_uniform(r,-1.,1.); p.z = reb_random_uniform(r,-1.,1.);
        p.vx = reb_random_normal(r, vel); p.vy = reb_random_normal(r, vel); p.vz = reb_random_normal(r, vel);
        p.m = 1./(double)n; p.r = radius;
        reb_simulation_add(r, p);
    }
}

static struct reb_simulation* make(unsigned int seed, int n){
    struct reb_simulation* r = reb_simulation_create();
    r->rand_seed = seed;
    reb_simulation_set_integrator(r, "leapfrog");
    r->softening = 0.05; r->dt = 0.01;
    cloud(r, n, 0.3, 0.);
    return r;
}

/* user callbacks of the "hooks" scenario */
static void drag(struct reb_simulation* r){                    /* additional_forces: reads velocities, adds to a */
    for (size_t i=0;i<r->N;i++){
        r->particles[i].ax -= 0.3*r->particles[i].vx; r->particles[i].ay -= 0.3*r->particles[i].vy; r->particles[i].az -= 0.3*r->particles[i].vz;
    }
}
static void pre_mod(struct reb_simulation* r){ r->particles[1].vz += 1e-3; }
static void post_mod(struct reb_simulation* r){ r->particles[2].m *= 1.0001; }
static int n_resolved = 0;
static enum REB_COLLISION_RESOLVE_OUTCOME eat(struct reb_simulation* const r, struct reb_collision c){
    /* the heavier particle swallows the other one's mass; the lighter one is removed */
    n_resolved++;
    struct reb_particle* a = &r->particles[c.p1]; struct reb_particle* b = &r->particles[c.p2];
    if (a->m >= b->m){ a->m += b->m; return REB_COLLISION_RESOLVE_OUTCOME_REMOVE_P2; }
    b->m += a->m; return REB_COLLISION_RESOLVE_OUTCOME_REMOVE_P1;
}

#define RND() (seed = seed*1664525u + 1013904223u, (seed>>8))
static int heartbeat_calls = 0;
static void heartbeat_count(struct reb_simulation* r){ (void)r; heartbeat_calls++; }

/* user ODE coupled to the particles ("odes" scenario): dy/dt = -y * |x_0|^2, reads r->particles every step */
static void ode_rhs(struct reb_ode* const ode, double* const yDot, const double* const y, const double t){
    (void)t;
    const struct reb_particle* p = &ode->r->particles[0];
    const double w = p->x*p->x + p->y*p->y + p->z*p->z;
    yDot[0] = -y[0]*w; yDot[1] = y[0]*p->vx;
}
static void ode_pre(struct reb_ode* const ode, const double* const y0){ (void)y0; ode->y[1] += 1e-3*ode->r->particles[1].x; }

/* heartbeats of the "lazy*" scenarios: what a heartbeat may do with r->particles */
static double hb_sum = 0.;
static void hb_blind(struct reb_simulation* r){ heartbeat_calls++; hb_sum += r->t; }                 /* never looks at the particles */
static void hb_reader(struct reb_simulation* r){                                                      /* reads them every 5th step */
    heartbeat_calls++;
    if (r->steps_done%5==0) for (size_t i=0;i<r->N;i+=97) hb_sum += r->particles[i].x + r->particles[i].vy;
}
static void hb_writer(struct reb_simulation* r){                                                      /* edits them every 7th step */
    heartbeat_calls++;
    if (r->steps_done%7==3){ r->particles[5].vx += 1e-3; r->particles[r->N-1].z *= 1.0001; }
}
static void hb_grower(struct reb_simulation* r){                                                      /* adds / removes particles */
    heartbeat_calls++;
    if (r->steps_done==4 || r->steps_done==9){
        struct reb_particle p = {0};
        p.x = 0.3 + 0.01*(double)r->steps_done; p.y = -0.2; p.z = 0.1; p.m = 1e-4;
        reb_simulation_add(r, p);
    }
    if (r->steps_done==6) reb_simulation_remove_particle(r, 17);
}

static void* thread_main(void* arg){
    struct reb_simulation* r = arg;
    reb_simulation_steps(r, 7);
    reb_simulation_integrate(r, r->t + 0.0555);
    return NULL;
}

int main(int argc, char** argv){
    if (argc<3){ fprintf(stderr, "usage: %s scenario outfile [N]\n", argv[0]); return 2; }
    const char* scen = argv[1];
    const int N = argc>3 ? atoi(argv[3]) : 100;
    out = fopen(argv[2], "wb");
    if (!out) return 3;
    if (strcmp(scen, "addremove")==0){
        /* the particle array grows (realloc may move it), shrinks and is edited between calls */
        struct reb_simulation* r = make(1, N);
        reb_simulation_steps(r, 3); dump("ar1 ", r);
        cloud(r, 3*N, 0.2, 0.);                          /* grows past N_allocated: r->particles moves */
        reb_simulation_steps(r, 3); dump("ar2 ", r);
        reb_simulation_remove_particle(r, 2);
        reb_simulation_remove_particle(r, r->N-1);
        reb_simulation_integrate(r, r->t + 0.0333); dump("ar3 ", r);
        r->particles[5].vx += 0.25; r->particles[0].m *= 2.;     /* raw edits after a synchronised call ... */
        r->did_modify_particles = 1;                             /* ... flagged as the reference asks for integrators that keep state (rebound.h:247) */
        reb_simulation_steps(r, 2); dump("ar4 ", r);
        reb_simulation_move_to_com(r);
        reb_simulation_steps(r, 2); dump("ar5 ", r);
        reb_simulation_free(r);
    }else if (strcmp(scen, "switch")==0){
        /* integrators and gravity modes change between calls */
        struct reb_simulation* r = make(2, N);
        reb_simulation_steps(r, 3); dump("sw1 ", r);
        reb_simulation_set_integrator(r, "ias15"); r->dt = 0.001;
        reb_simulation_steps(r, 2); dump("sw2 ", r);
        reb_simulation_set_integrator(r, "leapfrog"); r->dt = 0.01; r->gravity = REB_GRAVITY_COMPENSATED;
        reb_simulation_integrate(r, r->t + 0.047); dump("sw3 ", r);
        r->gravity = REB_GRAVITY_TREE; r->root_size = 40.; r->opening_angle2 = 0.3;
        reb_simulation_steps(r, 3); dump("sw4 ", r);
        r->gravity = REB_GRAVITY_NONE;
        reb_simulation_steps(r, 2); dump("sw5 ", r);
        r->gravity = REB_GRAVITY_BASIC; r->N_active = N/4; r->testparticle_type = 1;
        reb_simulation_steps(r, 3); dump("sw6 ", r);
        r->testparticle_type = 0;
        reb_simulation_integrate(r, r->t + 0.05); dump("sw7 ", r);
        reb_simulation_free(r);
    }else if (strcmp(scen, "copy")==0){
        /* copies and diffs of a simulation between calls; two simulations advance side by side */
        struct reb_simulation* r = make(3, N);
        reb_simulation_steps(r, 4);
        struct reb_simulation* r2 = reb_simulation_copy(r);
        reb_simulation_steps(r, 3);
        reb_simulation_steps(r2, 3);
        const int differ = reb_simulation_diff(r, r2, 2);
        fwrite(&differ, sizeof(int), 1, out);
        dump("cp1 ", r); dump("cp2 ", r2);
        struct reb_simulation* r3 = make(33, N/2);
        for (int k=0;k<4;k++){ reb_simulation_steps(r, 2); reb_simulation_steps(r3, 3); }
        dump("cp3 ", r); dump("cp4 ", r3);
        reb_simulation_free(r); reb_simulation_free(r2); reb_simulation_free(r3);
    }else if (strcmp(scen, "error")==0){
        /* tree gravity with a particle outside the box: the error ends an integration with REB_STATUS_GENERIC_ERROR */
        struct reb_simulation* r = make(4, N);
        r->save_messages = 1;
        r->gravity = REB_GRAVITY_TREE; r->root_size = 4.; r->opening_angle2 = 0.3;
        reb_simulation_steps(r, 2); dump("er1 ", r);
        r->particles[7].x = 9.;                          /* outside */
        r->did_modify_particles = 1;
        enum REB_STATUS st = reb_simulation_integrate(r, r->t + 0.2);
        const int sti = (int)st;
        fwrite(&sti, sizeof(int), 1, out);
        int n_err = 0;
        if (r->messages) for (int i=0;i<10;i++) if (r->messages[i] && r->messages[i][0]=='e') n_err++;
        const int has_err = n_err>0;
        fwrite(&has_err, sizeof(int), 1, out);
        double hdr[2] = {(double)r->N, (double)r->status};
        fwrite(hdr, sizeof(double), 2, out);
        reb_simulation_free(r);
    }else if (strcmp(scen, "short")==0){
        /* integrations shorter than a step, to the current time, and of exactly a few steps */
        struct reb_simulation* r = make(5, N);
        reb_simulation_integrate(r, r->t); dump("sh1 ", r);
        reb_simulation_integrate(r, 0.0042); dump("sh2 ", r);
        reb_simulation_integrate(r, r->t + 3.*r->dt); dump("sh3 ", r);
        reb_simulation_integrate(r, r->t + 5.*r->dt); dump("sh4 ", r);
        r->exact_finish_time = 0;
        reb_simulation_integrate(r, r->t + 4.5*r->dt); dump("sh5 ", r);
        reb_simulation_integrate(r, r->t - 6.2*fabs(r->dt)); dump("sh6 ", r);
        reb_simulation_steps(r, 1); dump("sh7 ", r);
        reb_simulation_free(r);
    }else if (strcmp(scen, "threads")==0){
        /* independent simulations stepped from several threads */
        enum { NT = 4 };
        struct reb_simulation* sims[NT]; pthread_t th[NT];
        for (int k=0;k<NT;k++) sims[k] = make(10+k, N + 7*k);
        for (int k=0;k<NT;k++) pthread_create(&th[k], NULL, thread_main, sims[k]);
        for (int k=0;k<NT;k++) pthread_join(th[k], NULL);
        for (int k=0;k<NT;k++){ dump("thr ", sims[k]); reb_simulation_free(sims[k]); }
    }else if (strcmp(scen, "hooks")==0){
        /* host callbacks between and inside the steps: additional_forces (the reference's host step runs, only its
         * force call is replaced), pre/post_timestep_modifications, a collision_resolve callback that removes
         * particles, an open boundary that tracks the energy of lost particles, variational particles (MEGNO) */
        struct reb_simulation* r = make(6, N);
        r->additional_forces = drag; r->force_is_velocity_dependent = 1;
        reb_simulation_steps(r, 4); dump("hk1 ", r);
        r->additional_forces = NULL;
        r->pre_timestep_modifications = pre_mod; r->post_timestep_modifications = post_mod;
        reb_simulation_integrate(r, r->t + 0.045); dump("hk2 ", r);
        r->pre_timestep_modifications = NULL; r->post_timestep_modifications = NULL;
        for (size_t i=0;i<r->N;i++) r->particles[i].r = 0.09;
        r->did_modify_particles = 1;
        r->collision = REB_COLLISION_DIRECT; r->collision_resolve = eat;
        reb_simulation_steps(r, 6); dump("hk3 ", r);
        fwrite(&n_resolved, sizeof(int), 1, out);
        r->collision = REB_COLLISION_NONE;
        r->boundary = REB_BOUNDARY_OPEN; r->root_size = 2.2; r->track_energy_offset = 1;
        for (size_t i=0;i<r->N;i+=3){ r->particles[i].vx *= 8.; r->particles[i].vy *= 8.; }
        r->did_modify_particles = 1;
        reb_simulation_steps(r, 25); dump("hk4 ", r);
        fwrite(&r->energy_offset, sizeof(double), 1, out);
        r->gravity = REB_GRAVITY_TREE; r->opening_angle2 = 0.3;
        reb_simulation_steps(r, 10); dump("hk5 ", r);
        fwrite(&r->energy_offset, sizeof(double), 1, out);
        reb_simulation_free(r);
        struct reb_simulation* v = make(7, 12);
        reb_simulation_init_megno_seed(v, 99);
        reb_simulation_integrate(v, 0.08);
        const double megno = reb_simulation_megno(v);
        fwrite(&megno, sizeof(double), 1, out);
        dump("hk6 ", v);
        reb_simulation_free(v);
    }else if (strcmp(scen, "seicache")==0){
        /* SEI caches sin/tan of OMEGA*dt in its state and refreshes them only when dt changes (integrator_sei.c:91-101):
         * the cache must look the same after device steps, and a changed OMEGA with an unchanged dt keeps the OLD
         * rotation in the reference -- the drop-in has to reproduce that, too */
        struct reb_simulation* r = reb_simulation_create();
        r->rand_seed = 8;
        reb_simulation_set_integrator(r, "sei");
        r->OMEGA = 1.0; r->dt = 1e-2; r->softening = 0.05; r->G = 1e-3;
        cloud(r, N, 0.1, 0.);
        reb_simulation_steps(r, 3); dump("se1 ", r);
        fwrite(r->integrator.state, sizeof(double), 5, out);
        fwrite(&r->OMEGAZ, sizeof(double), 1, out);
        r->OMEGA = 1.7;                                    /* dt unchanged: stale cache */
        reb_simulation_steps(r, 3); dump("se2 ", r);
        fwrite(r->integrator.state, sizeof(double), 5, out);
        r->dt = 2e-2;                                      /* refresh */
        reb_simulation_integrate(r, r->t + 0.11); dump("se3 ", r);
        fwrite(r->integrator.state, sizeof(double), 5, out);
        r->OMEGAZ = 2.5;
        reb_simulation_steps(r, 2); dump("se4 ", r);
        fwrite(r->integrator.state, sizeof(double), 5, out);
        r->OMEGAZ = -1;
        reb_simulation_steps(r, 2); dump("se5 ", r);
        fwrite(r->integrator.state, sizeof(double), 5, out);
        fwrite(&r->OMEGAZ, sizeof(double), 1, out);
        reb_simulation_free(r);
    }else if (strncmp(scen, "fuzs", 4)==0){
        /* the same idea for the shearing-sheet family: SEI, shear boundary, tree / direct gravity and collisions with the
         * reference's hard-sphere resolver, ghost boxes; the sequence also changes dt (SEI refreshes its cache) and OMEGA
         * (it does not), the ghost rings and the root-box grid */
        unsigned int seed = (unsigned int)atoi(scen+4) * 2246822519u + 777u;
        struct reb_simulation* r = reb_simulation_create();
        r->rand_seed = 70 + (unsigned int)atoi(scen+4);
        reb_simulation_set_integrator(r, "sei");
        r->boundary = REB_BOUNDARY_SHEAR; r->gravity = REB_GRAVITY_TREE; r->opening_angle2 = 0.5;
        r->OMEGA = 1.0; r->G = 2e-3; r->softening = 0.02; r->dt = 2e-2;
        r->root_size = 4.; r->N_root_x = 2; r->N_root_y = 2; r->N_ghost_x = 1; r->N_ghost_y = 1;
        r->collision_resolve = reb_collision_resolve_hardsphere; r->collision = REB_COLLISION_TREE;
        for (int i=0;i<N;i++){
            struct reb_particle p = {0};
            p.x = reb_random_uniform(r,-3.9,3.9); p.y = reb_random_uniform(r,-3.9,3.9); p.z = reb_random_normal(r, 0.05);
            p.vy = -1.5*p.x*r->OMEGA; p.vx = reb_random_normal(r, 0.02);
            p.r = reb_random_uniform(r, 0.15, 0.3); p.m = p.r*p.r*p.r;
            reb_simulation_add(r, p);
        }
        for (int op=0; op<50; op++){
            const unsigned int k = RND()%12;
            switch (k){
                case 0: case 1: case 2: reb_simulation_steps(r, 1 + RND()%5); break;
                case 3: reb_simulation_integrate(r, r->t + (0.4 + (RND()%500)/100.)*r->dt); break;
                case 4: { struct reb_particle p = {0};
                          p.x = reb_random_uniform(r,-3.9,3.9); p.y = reb_random_uniform(r,-3.9,3.9); p.z = 0.3 + 0.01*op;
                          p.vy = -1.5*p.x*r->OMEGA; p.r = 0.06; p.m = 2e-4; reb_simulation_add(r, p); break; }
                case 5: if (r->N > 10) reb_simulation_remove_particle(r, RND()%r->N); break;
                case 6: { const unsigned int c = RND()%3;
                          r->collision = c==0 ? REB_COLLISION_TREE : c==1 ? REB_COLLISION_DIRECT : REB_COLLISION_NONE; break; }
                case 7: r->dt = (RND()%2) ? 2e-2 : 1.5e-2; break;                      /* the SEI cache is refreshed */
                case 8: r->OMEGA = (RND()%2) ? 1.0 : 1.25; break;                      /* ... and here it is not */
                case 9: r->N_ghost_x = r->N_ghost_y = 1 + (int)(RND()%2); break;
                case 10: { struct reb_simulation* c2 = reb_simulation_copy(r); reb_simulation_free(r); r = c2; break; }
                /* (no REB_GRAVITY_BASIC here: with ghost boxes the reference's serial build applies the shifted pairs
                 * antisymmetrically and differs from its own OpenMP build, and from the gather form, in the last bit) */
                case 11: r->gravity = (RND()%3==0) ? REB_GRAVITY_NONE : REB_GRAVITY_TREE; break;
            }
            double hdr[5] = {(double)k, (double)r->N, r->t, (double)r->steps_done, (double)r->collisions_log_n};
            fwrite(hdr, sizeof(double), 5, out);
        }
        dump("fzs ", r);
        fwrite(r->integrator.state, sizeof(double), 5, out);
        reb_simulation_free(r);
    }else if (strncmp(scen, "fuzz", 4)==0){
        /* a pseudo-random sequence of public-API calls (seed = the digits after "fuzz"): stepping and integrating in
         * both directions, adding / removing / editing particles, switching gravity, integrator, collisions, boundary,
         * heartbeat, exit distance and test-particle settings, copying the simulation, reading diagnostics */
        unsigned int seed = (unsigned int)atoi(scen+4) * 2654435761u + 12345u;
        struct reb_simulation* r = make(50 + (unsigned int)atoi(scen+4), N);
        r->root_size = 60.;
        int hb_on = 0;
        for (int op=0; op<60; op++){
            const unsigned int k = RND()%15;
            switch (k){
                case 0: case 1: reb_simulation_steps(r, 1 + RND()%6); break;
                case 2: { const double f = 0.3 + (RND()%740)/100.; const double sgn = (RND()%5==0) ? -1. : 1.;
                          reb_simulation_integrate(r, r->t + sgn*f*fabs(r->dt)); break; }
                case 3: cloud(r, 1 + RND()%4, 0.2, r->collision ? 0.05 : 0.); break;
                case 4: if (r->N > 12) reb_simulation_remove_particle(r, RND()%r->N); break;
                case 5: if (r->N){ r->particles[RND()%r->N].vy += 0.05; r->did_modify_particles = 1; } break;
                case 6: { const unsigned int g = RND()%4;
                          r->gravity = g==0 ? REB_GRAVITY_BASIC : g==1 ? REB_GRAVITY_COMPENSATED : g==2 ? REB_GRAVITY_TREE : REB_GRAVITY_NONE; break; }
                case 7: if (RND()%2){ reb_simulation_set_integrator(r, "ias15"); r->dt = 2e-3; }
                        else { reb_simulation_set_integrator(r, "leapfrog"); r->dt = copysign(0.01, r->dt); } break;
                case 8: hb_on = !hb_on; r->heartbeat = hb_on ? heartbeat_count : NULL; break;
                case 9: r->exit_max_distance = r->exit_max_distance ? 0. : 25.; break;
                case 10: if (RND()%2){ r->N_active = r->N/2; r->testparticle_type = (int)(RND()%2); } else { r->N_active = SIZE_MAX; } break;
                case 11: { struct reb_simulation* c2 = reb_simulation_copy(r); reb_simulation_free(r); r = c2;
                           r->heartbeat = hb_on ? heartbeat_count : NULL; break; }
                case 12: reb_simulation_move_to_com(r); break;
                case 13: { const double e = reb_simulation_energy(r); fwrite(&e, sizeof(double), 1, out); break; }
                case 14: if (r->collision){ r->collision = REB_COLLISION_NONE; }
                         else { r->collision = (RND()%2) ? REB_COLLISION_DIRECT : REB_COLLISION_TREE; r->collision_resolve = reb_collision_resolve_merge;
                                for (size_t i=0;i<r->N;i++) r->particles[i].r = 0.03; r->did_modify_particles = 1; } break;
            }
            if (r->status > 0 && r->status != REB_STATUS_SUCCESS) r->status = REB_STATUS_SUCCESS;     /* an escape ends one call, not the sequence */
            double hdr[4] = {(double)k, (double)r->N, r->t, (double)r->steps_done};
            fwrite(hdr, sizeof(double), 4, out);
        }
        dump("fuz ", r);
        fwrite(&heartbeat_calls, sizeof(int), 1, out);
        reb_simulation_free(r);
    }else if (strcmp(scen, "odes")==0){
        /* a user ODE whose right-hand side and pre_timestep hook read r->particles: integrated on the host after every
         * N-body step (simulation.c:531-556), so the host copy must be current at every step */
        struct reb_simulation* r = make(21, N);
        struct reb_ode* ode = reb_ode_create(r, 2);
        ode->derivatives = ode_rhs; ode->pre_timestep = ode_pre; ode->needs_nbody = 0;
        ode->y[0] = 1.0; ode->y[1] = 0.0;
        reb_simulation_steps(r, 6); dump("od1 ", r); fwrite(ode->y, sizeof(double), 2, out);
        reb_simulation_integrate(r, r->t + 0.0777); dump("od2 ", r); fwrite(ode->y, sizeof(double), 2, out);
        reb_ode_free(ode);
        reb_simulation_steps(r, 5); dump("od3 ", r);
        reb_simulation_free(r);
    }else if (strncmp(scen, "lazy", 4)==0){
        /* a heartbeat is installed (as in every example of the reference): the particle array is large enough for
         * the drop-in to keep it on the device and fetch the host copy only when the heartbeat touches it */
        const int n = N < 3000 ? 3000 : N;
        struct reb_simulation* r = make(31, n);
        hb_sum = 0.; heartbeat_calls = 0;
        r->heartbeat = strcmp(scen,"lazy_blind")==0 ? hb_blind : strcmp(scen,"lazy_read")==0 ? hb_reader
                     : strcmp(scen,"lazy_write")==0 ? hb_writer : hb_grower;
        reb_simulation_steps(r, 12); dump("lz1 ", r);
        reb_simulation_integrate(r, r->t + 8.5*r->dt); dump("lz2 ", r);
        fwrite(&hb_sum, sizeof(double), 1, out); fwrite(&heartbeat_calls, sizeof(int), 1, out);
        r->heartbeat = NULL;
        reb_simulation_steps(r, 3); dump("lz3 ", r);
        reb_simulation_free(r);
    }else if (strcmp(scen, "sigint")==0){
        /* Ctrl-C during a long integration (the test harness makes the mock engine raise SIGINT in the middle of a
         * device batch): the run must end there with REB_STATUS_SIGINT, synchronised, short of tmax */
        struct reb_simulation* r = make(22, N);
        const enum REB_STATUS st = reb_simulation_integrate(r, 50000.*r->dt);
        double info[4] = {(double)st, (double)r->status, r->t, (double)r->steps_done};
        fwrite(info, sizeof(double), 4, out);
        dump("sig ", r);
        reb_simulation_free(r);
    }else if (strcmp(scen, "many")==0){
        /* a parameter sweep: hundreds of short-lived simulations, created and freed one after the other (and a few
         * kept alive), each using the replaced hot path */
        struct reb_simulation* keep[8] = {0};
        for (int k=0;k<700;k++){
            struct reb_simulation* r = make(100+k, 8 + k%5);
            r->save_messages = 1;
            reb_simulation_steps(r, 2);
            int n_err = 0;
            if (r->messages) for (int i=0;i<10;i++) if (r->messages[i] && r->messages[i][0]=='e') n_err++;
            if (n_err){ fprintf(stderr, "simulation %d: error message queued\n", k); return 5; }
            if (k%97==0) dump("mny ", r);
            if (k%100==3 && k/100<8) keep[k/100] = r; else reb_simulation_free(r);
        }
        for (int k=0;k<8;k++) if (keep[k]){ reb_simulation_steps(keep[k], 1); dump("kep ", keep[k]); reb_simulation_free(keep[k]); }
        /* ... and a list of simulations that are all alive at once, advanced round-robin */
        enum { NLIVE = 600 };
        static struct reb_simulation* live[NLIVE];
        for (int k=0;k<NLIVE;k++){ live[k] = make(1000+k, 6 + k%4); live[k]->save_messages = 1; }
        for (int round=0;round<2;round++) for (int k=0;k<NLIVE;k++) reb_simulation_steps(live[k], 2);
        for (int k=0;k<NLIVE;k++){
            int n_err = 0;
            if (live[k]->messages) for (int i=0;i<10;i++) if (live[k]->messages[i] && live[k]->messages[i][0]=='e') n_err++;
            if (n_err){ fprintf(stderr, "live simulation %d: error message queued\n", k); return 6; }
            if (k%59==0) dump("liv ", live[k]);
            reb_simulation_free(live[k]);
        }
    }else{ fprintf(stderr, "unknown scenario %s\n", scen); return 2; }
    fclose(out);
    return 0;
}
