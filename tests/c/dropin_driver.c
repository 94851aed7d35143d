/*
 * dropin_driver.c -- drives a simulation purely through the reference's public C API
 * (reb_simulation_create / reb_simulation_add* / reb_simulation_set_integrator / reb_simulation_steps)
 * and dumps the final state.  The SAME source is linked once against the unmodified reference and once
 * against the drop-in librebound (reference sources + CUDA hot path); tests/test_gpu_dropin.py compares
 * the two dumps bit for bit.  Scenarios follow the reference's examples.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "rebound.h"
#include "integrator_leapfrog.h"
#include "integrator_eos.h"

static double restitution_bridges(const struct reb_simulation* const r, double v){
    (void)r;                                   /* examples/shearing_sheet/problem.c:96-103 */
    double eps = 0.32*pow(fabs(v)*100.,-0.234);
    if (eps>1) eps=1;
    if (eps<0) eps=0;
    return eps;
}

static int heartbeat_calls = 0;
static void heartbeat(struct reb_simulation* r){ (void)r; heartbeat_calls++; }

/* "archive" scenario: diagnostics read in the middle of a run, while a resident simulation is unsynchronised */
static double mid_energy = 0., mid_com_x = 0., mid_Lz = 0.;
static void heartbeat_mid(struct reb_simulation* r){
    heartbeat_calls++;
    if (r->steps_done==3){
        mid_energy = reb_simulation_energy(r);
        mid_com_x = reb_simulation_com(r).x;
        mid_Lz = reb_simulation_angular_momentum(r).z;
    }
}

int main(int argc, char** argv){
    if (argc<3){ fprintf(stderr, "usage: %s scenario outfile [N] [steps]\n", argv[0]); return 2; }
    const char* scen = argv[1];
    int N = argc>3 ? atoi(argv[3]) : 1000;
    int steps = argc>4 ? atoi(argv[4]) : 5;
    struct reb_simulation* r = reb_simulation_create();
    r->rand_seed = 42;
    char sa_file[4096] = {0};
    int use_integrate = 0;      /* reb_simulation_integrate(r, steps*dt) instead of reb_simulation_steps(r, steps) */
    double integrate_sign = 1.;
    if (strcmp(scen, "plummer")==0 || strcmp(scen, "plummer_comp")==0 || strcmp(scen, "archive")==0 || strcmp(scen, "edit")==0){
        /* examples/selfgravity_plummer/problem.c */
        double M=1, R=1, E=3./64.*M_PI*M*M/R, r0=16./(3.*M_PI)*R;
        double t0 = r->G*pow(M,5./2.)*pow(4.*E,-3./2.)*(double)N/log(0.4*(double)N);
        reb_simulation_set_integrator(r, "leapfrog");
        r->dt = 2e-5*t0; r->softening = 0.01*r0;
        if (strcmp(scen, "plummer_comp")==0) r->gravity = REB_GRAVITY_COMPENSATED;
        reb_simulation_add_plummer(r, N, M, R);
        reb_simulation_move_to_com(r);
        r->heartbeat = heartbeat;
        if (strcmp(scen, "edit")==0) r->heartbeat = NULL;      /* nothing observes the particles between steps */
        if (strcmp(scen, "archive")==0){
            /* Simulationarchive snapshot every 2 steps, written from inside reb_simulation_steps */
            snprintf(sa_file, sizeof(sa_file), "%s.sa", argv[2]);
            remove(sa_file);
            reb_simulation_save_to_file_step(r, sa_file, 2);
            r->heartbeat = heartbeat_mid;
        }
    }else if (strcmp(scen, "testparticles")==0){
        reb_simulation_set_integrator(r, "leapfrog");
        r->dt = 1e-2;
        struct reb_particle star = {0}; star.m = 1; reb_simulation_add(r, star);
        for (int i=1;i<10;i++) reb_simulation_add_fmt(r, "m a", 1e-4*i, (double)i);
        r->N_active = 10; r->testparticle_type = 1;
        for (int i=0;i<N;i++) reb_simulation_add_fmt(r, "m a e omega f", 1e-9, reb_random_uniform(r,0.4,20.), reb_random_uniform(r,0.01,0.2),
                                                     reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,2.*M_PI));
    }else if (strcmp(scen, "disc")==0){
        /* examples/selfgravity_disc/problem.c */
        reb_simulation_set_integrator(r, "leapfrog");
        r->gravity = REB_GRAVITY_TREE; r->boundary = REB_BOUNDARY_OPEN; r->opening_angle2 = 0.25;
        r->G = 1; r->softening = 0.02; r->dt = 3e-2;
        const double boxsize = 10.2; r->root_size = boxsize;
        double disc_mass = 2e-1;
        struct reb_particle star = {0}; star.m = 1; reb_simulation_add(r, star);
        for (int i=0;i<N;i++){
            struct reb_particle pt = {0};
            double a = reb_random_powerlaw(r, boxsize/10.,boxsize/2./1.2,-1.5);
            double phi = reb_random_uniform(r, 0,2.*M_PI);
            pt.x = a*cos(phi); pt.y = a*sin(phi); pt.z = a*reb_random_normal(r, 0.001);
            double mu = star.m + disc_mass*(pow(a,-3./2.)-pow(boxsize/10.,-3./2.))/(pow(boxsize/2./1.2,-3./2.)-pow(boxsize/10.,-3./2.));
            double vkep = sqrt(r->G*mu/a);
            pt.vx = vkep*sin(phi); pt.vy = -vkep*cos(phi); pt.m = disc_mass/(double)N;
            reb_simulation_add(r, pt);
        }
    }else if (strcmp(scen, "sheet")==0 || strcmp(scen, "sheet_hb")==0){
        /* examples/shearing_sheet/problem.c; N is the root box size in metres.  sheet_hb: with a heartbeat installed, as
         * in the example (it prints the time and does not touch the particles) */
        if (strcmp(scen, "sheet_hb")==0) r->heartbeat = heartbeat;
        r->opening_angle2 = .5;
        reb_simulation_set_integrator(r, "sei");
        r->boundary = REB_BOUNDARY_SHEAR; r->gravity = REB_GRAVITY_TREE; r->collision = REB_COLLISION_TREE;
        r->collision_resolve = reb_collision_resolve_hardsphere;
        r->OMEGA = 0.00013143527; r->G = 6.67428e-11; r->softening = 0.1;
        r->dt = 1e-3*2.*M_PI/r->OMEGA;
        r->root_size = (double)N; r->N_root_x = 2; r->N_root_y = 2; r->N_ghost_x = 2; r->N_ghost_y = 2; r->N_ghost_z = 0;
        double bx = r->root_size*2., by = r->root_size*2.;
        r->coefficient_of_restitution = restitution_bridges;
        r->minimum_collision_velocity = 1.*r->OMEGA*0.001;
        double total_mass = 400.*bx*by, mass = 0;
        while (mass<total_mass){
            struct reb_particle pt = {0};
            pt.x = reb_random_uniform(r, -bx/2.,bx/2.); pt.y = reb_random_uniform(r, -by/2.,by/2.); pt.z = reb_random_normal(r, 1.);
            pt.vy = -1.5*pt.x*r->OMEGA;
            double radius = reb_random_powerlaw(r, 1., 4., -3.);
            pt.r = radius; pt.m = 400.*4./3.*M_PI*radius*radius*radius;
            reb_simulation_add(r, pt);
            mass += pt.m;
        }
    }else if (strncmp(scen, "lf", 2)==0 && strlen(scen)==3){
        /* higher-order leapfrog (integrator_leapfrog.c:101-208): lf4, lf6, lf8 */
        double M=1, R=1, E=3./64.*M_PI*M*M/R, r0=16./(3.*M_PI)*R;
        double t0 = r->G*pow(M,5./2.)*pow(4.*E,-3./2.)*(double)N/log(0.4*(double)N);
        reb_simulation_set_integrator(r, "leapfrog");
        ((struct reb_integrator_leapfrog_state*)r->integrator.state)->order = (unsigned int)(scen[2]-'0');
        r->dt = 2e-4*t0; r->softening = 0.01*r0;
        reb_simulation_add_plummer(r, N, M, R);
    }else if (strncmp(scen, "ias15", 5)==0 || strcmp(scen, "whfast")==0){
        /* integrators outside the GPU hot path that call reb_simulation_update_acceleration: IAS15 with
         * BASIC or COMPENSATED gravity (reads r->gravity_cs, integrator_ias15.c:337-343), WHFast (Jacobi terms stay on
         * the reference path, the rest goes through reb_gravity_basic with gravity_ignore_terms) */
        struct reb_particle star = {0}; star.m = 1; reb_simulation_add(r, star);
        for (int i=0;i<N;i++) reb_simulation_add_fmt(r, "m a e omega f inc", 1e-6, reb_random_uniform(r,1.,30.), reb_random_uniform(r,0.0,0.1),
                                                     reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,0.05));
        reb_simulation_move_to_com(r);
        if (strcmp(scen, "whfast")==0){ reb_simulation_set_integrator(r, "whfast"); r->dt = 0.05; }
        else { reb_simulation_set_integrator(r, "ias15"); r->dt = 0.01; if (strcmp(scen, "ias15_comp")==0) r->gravity = REB_GRAVITY_COMPENSATED; }
    }else if (strcmp(scen, "tp0")==0){
        /* examples/solar_system_with_testparticles: massless planetesimals, testparticle_type 0 */
        reb_simulation_set_integrator(r, "leapfrog");
        r->dt = 1e-2;
        struct reb_particle star = {0}; star.m = 1; reb_simulation_add(r, star);
        for (int i=1;i<10;i++) reb_simulation_add_fmt(r, "m a", 1e-4*i, (double)i);
        r->N_active = 10; r->testparticle_type = 0;
        for (int i=0;i<N;i++){      /* explicit primary: reb_simulation_add_fmt would recompute the centre of mass per particle */
            const double a = reb_random_uniform(r,0.4,20.), e = reb_random_uniform(r,0.01,0.2);
            const double omega = reb_random_uniform(r,0.,2.*M_PI), f = reb_random_uniform(r,0.,2.*M_PI);
            reb_simulation_add(r, reb_particle_from_orbit(r->G, r->particles[0], 0., a, e, 0., 0., omega, f));
        }
    }else if (strcmp(scen, "merge")==0 || strcmp(scen, "line")==0){
        /* a cold cloud of big particles: DIRECT search + merging (collision.c:64-124, 674-737), or the LINE search
         * with hard-sphere bounces (collision.c:125-196) */
        reb_simulation_set_integrator(r, "leapfrog");
        r->gravity = REB_GRAVITY_BASIC; r->dt = 2e-2; r->softening = 0.01;
        if (strcmp(scen, "merge")==0){ r->collision = REB_COLLISION_DIRECT; r->collision_resolve = reb_collision_resolve_merge; }
        else { r->collision = REB_COLLISION_LINE; r->collision_resolve = reb_collision_resolve_hardsphere; }
        for (int i=0;i<N;i++){
            struct reb_particle pt = {0};
            pt.x = reb_random_uniform(r,-1.,1.); pt.y = reb_random_uniform(r,-1.,1.); pt.z = reb_random_uniform(r,-1.,1.);
            pt.vx = reb_random_normal(r, 0.05); pt.vy = reb_random_normal(r, 0.05); pt.vz = reb_random_normal(r, 0.05);
            pt.m = 1./(double)N; pt.r = 0.4*pow((double)N, -1./3.)*reb_random_uniform(r, 0.5, 1.);
            reb_simulation_add(r, pt);
        }
    }else if (strcmp(scen, "periodic")==0){
        /* tree gravity in a periodic box with one ring of ghost boxes, 2x2x1 root boxes */
        reb_simulation_set_integrator(r, "leapfrog");
        r->gravity = REB_GRAVITY_TREE; r->boundary = REB_BOUNDARY_PERIODIC; r->opening_angle2 = 0.3;
        r->softening = 0.05; r->dt = 5e-2;
        r->root_size = 4.; r->N_root_x = 2; r->N_root_y = 2; r->N_root_z = 1;
        r->N_ghost_x = 1; r->N_ghost_y = 1; r->N_ghost_z = 1;
        for (int i=0;i<N;i++){
            struct reb_particle pt = {0};
            pt.x = reb_random_uniform(r,-4.,4.); pt.y = reb_random_uniform(r,-4.,4.); pt.z = reb_random_uniform(r,-2.,2.);
            pt.vx = reb_random_normal(r, 1.); pt.vy = reb_random_normal(r, 1.); pt.vz = reb_random_normal(r, 1.);
            pt.m = 1./(double)N;
            reb_simulation_add(r, pt);
        }
    }else if (strcmp(scen, "open_direct")==0){
        /* direct gravity in an open box: fast particles leave and are removed (boundary.c:44-77) */
        reb_simulation_set_integrator(r, "leapfrog");
        r->gravity = REB_GRAVITY_BASIC; r->boundary = REB_BOUNDARY_OPEN; r->softening = 0.05; r->dt = 5e-2;
        r->root_size = 6.;
        for (int i=0;i<N;i++){
            struct reb_particle pt = {0};
            pt.x = reb_random_uniform(r,-2.,2.); pt.y = reb_random_uniform(r,-2.,2.); pt.z = reb_random_uniform(r,-2.,2.);
            pt.vx = reb_random_normal(r, 3.); pt.vy = reb_random_normal(r, 3.); pt.vz = reb_random_normal(r, 3.);
            pt.m = 1./(double)N;
            reb_simulation_add(r, pt);
        }
    }else if (strcmp(scen, "mercurius")==0 || strcmp(scen, "trace")==0){
        /* hybrid integrators: close encounters are integrated for a subset of the particles, and the collision
         * search runs on that subset through r->map / r->N_map (integrator_mercurius.c:404-429, integrator_trace.c:820)
         * and, between steps, against the star only through r->N_targets = 1 (integrator_mercurius.c:909) */
        struct reb_particle star = {0}; star.m = 1; star.r = 0.005; reb_simulation_add(r, star);
        for (int i=0;i<N;i++) reb_simulation_add_fmt(r, "m a e omega f inc r", 3e-4, reb_random_uniform(r,1.,1.6), reb_random_uniform(r,0.0,0.3),
                                                     reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,0.02), 0.004);
        reb_simulation_move_to_com(r);
        reb_simulation_set_integrator(r, scen);
        r->dt = 0.02;
        r->collision = REB_COLLISION_DIRECT; r->collision_resolve = reb_collision_resolve_merge;
    }else if (strncmp(scen, "integ_", 6)==0){
        /* reb_simulation_integrate on a simulation nothing observes between steps (the drop-in batches the whole steps
         * on the device and leaves the end of the run to the reference's exit logic): integ_exact ends on tmax with a
         * shortened last step (exact_finish_time=1), integ_over steps past it, integ_back runs backwards in time,
         * integ_tree is the disc with tree gravity and an open boundary */
        reb_simulation_set_integrator(r, "leapfrog");
        if (strcmp(scen, "integ_tree")==0){
            r->gravity = REB_GRAVITY_TREE; r->boundary = REB_BOUNDARY_OPEN; r->opening_angle2 = 0.25;
            r->softening = 0.02; r->dt = 3e-2; r->root_size = 10.2;
            struct reb_particle star = {0}; star.m = 1; reb_simulation_add(r, star);
            for (int i=0;i<N;i++){
                struct reb_particle pt = {0};
                double a = reb_random_powerlaw(r, 1.02, 4.25, -1.5), phi = reb_random_uniform(r, 0,2.*M_PI);
                pt.x = a*cos(phi); pt.y = a*sin(phi); pt.z = a*reb_random_normal(r, 0.001);
                double vkep = sqrt(1.1/a)*reb_random_uniform(r, 0.9, 3.0);       /* some particles leave the box */
                pt.vx = vkep*sin(phi); pt.vy = -vkep*cos(phi); pt.m = 0.2/(double)N;
                reb_simulation_add(r, pt);
            }
        }else{
            double M=1, R=1, E=3./64.*M_PI*M*M/R, r0=16./(3.*M_PI)*R;
            double t0 = r->G*pow(M,5./2.)*pow(4.*E,-3./2.)*(double)N/log(0.4*(double)N);
            r->dt = 2e-5*t0; r->softening = 0.01*r0;
            reb_simulation_add_plummer(r, N, M, R);
        }
        use_integrate = 2;
        r->exact_finish_time = strcmp(scen, "integ_over")==0 ? 0 : 1;
        if (strcmp(scen, "integ_back")==0) integrate_sign = -1.;
    }else if (strcmp(scen, "eos")==0){
        /* Embedded operator splitting with a modified-kick outer scheme (PMLF4): every interaction step is a force
         * evaluation followed by reb_gravity_basic_calculate_and_apply_jerk (integrator_eos.c:96-110) */
        struct reb_particle star = {0}; star.m = 1; reb_simulation_add(r, star);
        for (int i=0;i<N;i++) reb_simulation_add_fmt(r, "m a e omega f inc", 1e-5, reb_random_uniform(r,1.,10.), reb_random_uniform(r,0.0,0.1),
                                                     reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,2.*M_PI), reb_random_uniform(r,0.,0.05));
        reb_simulation_move_to_com(r);
        struct reb_integrator_eos_state* eos = reb_simulation_set_integrator(r, "eos");
        eos->phi0 = REB_INTEGRATOR_EOS_TYPE_PMLF4; eos->phi1 = REB_INTEGRATOR_EOS_TYPE_LF4; eos->n = 2;
        r->dt = 0.05;
    }else if (strcmp(scen, "escape")==0 || strcmp(scen, "encounter")==0){
        /* run_heartbeat's exit conditions (simulation.c:242-272): a hot cloud loses a particle past exit_max_distance,
         * or two particles come closer than exit_min_distance; reb_simulation_integrate ends with that status */
        reb_simulation_set_integrator(r, "leapfrog");
        r->gravity = REB_GRAVITY_BASIC; r->softening = 0.05; r->dt = 2e-2;
        for (int i=0;i<N;i++){
            struct reb_particle pt = {0};
            pt.x = reb_random_uniform(r,-1.,1.); pt.y = reb_random_uniform(r,-1.,1.); pt.z = reb_random_uniform(r,-1.,1.);
            pt.vx = reb_random_normal(r, 1.); pt.vy = reb_random_normal(r, 1.); pt.vz = reb_random_normal(r, 1.);
            pt.m = 1./(double)N;
            reb_simulation_add(r, pt);
        }
        if (strcmp(scen, "escape")==0) r->exit_max_distance = 4.; else r->exit_min_distance = 1e-2;
        use_integrate = 1;
    }else{ fprintf(stderr, "unknown scenario %s\n", scen); return 2; }

    struct timespec t_begin, t_end;
    if (getenv("DRIVER_WARMUP")) reb_simulation_steps(r, 1);      /* timing runs: CUDA context + first upload outside the clock */
    clock_gettime(CLOCK_MONOTONIC, &t_begin);
    if (use_integrate==2){
        /* a final time that is not a whole number of steps away, reached in two calls */
        reb_simulation_integrate(r, r->t + integrate_sign*(0.4*steps + 0.37)*r->dt);
        reb_simulation_integrate(r, r->t + integrate_sign*(0.6*steps + 0.21)*fabs(r->dt));
    }else if (use_integrate){ r->exact_finish_time = 0; reb_simulation_integrate(r, r->t + steps*r->dt); }
    else reb_simulation_steps(r, steps);
    clock_gettime(CLOCK_MONOTONIC, &t_end);
    if (strcmp(scen, "edit")==0){
        /* plain host edits between two calls, without r->did_modify_particles: legal with the reference's leapfrog */
        for (size_t i=0;i<r->N;i+=7){ r->particles[i].vx += 0.125; r->particles[i].y *= 1.5; }
        reb_simulation_steps(r, steps);
    }

    FILE* f = fopen(argv[2], "wb");
    if (!f) return 3;
    double hdr[6] = {(double)r->N, r->t, (double)r->collisions_log_n, r->collisions_plog, (double)r->status, (double)heartbeat_calls};
    fwrite(hdr, sizeof(double), 6, f);
    for (size_t i=0;i<r->N;i++) fwrite(&r->particles[i], sizeof(double), 11, f);
    if (sa_file[0]){
        /* restart from the snapshot taken after 2 steps and run to the same step count: must land on the same bits
         * (python_tests/test_simulationarchive.py:610-630) */
        struct reb_simulation* r2 = reb_simulation_create_from_file(sa_file, 1);
        if (!r2) return 4;
        const double restart_steps_done = (double)r2->steps_done;
        reb_simulation_steps(r2, (size_t)steps - (size_t)r2->steps_done);
        double tail[6] = {(double)r2->N, r2->t, restart_steps_done, mid_energy, mid_com_x, mid_Lz};
        fwrite(tail, sizeof(double), 6, f);
        for (size_t i=0;i<r2->N;i++) fwrite(&r2->particles[i], sizeof(double), 11, f);
        reb_simulation_free(r2);
        remove(sa_file);
    }
    fclose(f);
    printf("%s N=%zu t=%.17g collisions=%lld steps_call_seconds=%.6f\n", scen, r->N, r->t, (long long)r->collisions_log_n,
           (t_end.tv_sec-t_begin.tv_sec) + 1e-9*(t_end.tv_nsec-t_begin.tv_nsec));
    reb_simulation_free(r);
    return 0;
}
