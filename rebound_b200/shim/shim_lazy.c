/*
 * shim_lazy.c -- r->particles stays on the device under a heartbeat: the host copy is fetched only when somebody
 * actually touches it.
 *
 * Every example the BASELINE configurations come from installs a heartbeat (examples/shearing_sheet/problem.c,
 * examples/selfgravity_disc/problem.c), and a heartbeat MAY read r->particles after every step -- so the automatic
 * residency used to give up and move the whole array both ways every step (117 MB each at N = 2^20).  Most heartbeats
 * only print the time or check an output interval.  Instead of guessing, the particle array is made inaccessible
 * while the device copy is ahead:
 *
 *   device step  ->  mprotect(r->particles, PROT_NONE)                      host copy stale, nobody can see it
 *   first read   ->  SIGSEGV -> download -> mprotect(PROT_READ) -> retry    host current, device still valid
 *   first write  ->  SIGSEGV -> mprotect(READ|WRITE), device copy dropped   the next step uploads again
 *
 * so a heartbeat that never looks at the particles costs nothing, one that reads them costs one download per read
 * step, and one that edits them costs what host-authoritative mode always cost.  Any other reader -- the exit-distance
 * scans of run_heartbeat, reb_simulation_add growing the array, user code between calls -- is caught the same way;
 * the end of the call (reb_simulation_synchronize) lifts the protection.
 *
 * What this needs:
 *   * the array must own whole pages: the reference allocates it with realloc (src/particle.c:56, src/binarydata.c);
 *     those translation units are compiled with -Drealloc=reb_b200_particles_realloc / -Dfree=reb_b200_particles_free
 *     (rebound_b200/shim/Makefile), which hand out page-aligned, page-rounded blocks for large requests and lift the
 *     protection (bringing the host copy up to date) before glibc copies or frees a protected block -- libc itself
 *     never touches a protected page, so the fault handler never runs inside malloc;
 *   * a SIGSEGV handler (chained to whatever was installed before) that recognises the protected ranges.  The fault is
 *     synchronous and comes from code that reads particles, never from inside the CUDA driver (the shim lifts the
 *     protection before every copy it issues itself), so calling the engine from the handler is safe.
 * REBOUND_B200_LAZY=0 switches the mechanism off (heartbeats then keep the simulation host-current as before).
 */
#include <malloc.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#include <pthread.h>
#include "shim_common.h"

#define LAZY_MAX 256
#define LAZY_MIN_BYTES (256u<<10)        /* smaller arrays are not worth a fault */
enum { LAZY_OFF = 0, LAZY_NONE = 1, LAZY_READ = 2 };      /* protection of a registered range */

struct lazy_range { char* lo; char* hi; struct shim_state* s; volatile int prot; };
static struct lazy_range ranges[LAZY_MAX];
static volatile int n_ranges = 0;
static pthread_mutex_t lazy_lock = PTHREAD_MUTEX_INITIALIZER;
static struct sigaction old_action;
static int handler_installed = 0;
static size_t page = 4096;

static int lazy_enabled(void){
    static int on = -1;
    if (on < 0){ const char* e = getenv("REBOUND_B200_LAZY"); on = (e && e[0]=='0') ? 0 : 1; page = (size_t)sysconf(_SC_PAGESIZE); }
    return on;
}

static struct lazy_range* find_addr(const void* addr){
    const char* a = (const char*)addr;
    for (int i=0;i<n_ranges;i++) if (ranges[i].prot!=LAZY_OFF && a>=ranges[i].lo && a<ranges[i].hi) return &ranges[i];
    return NULL;
}

/* Host copy current and fully accessible; the device copy stays valid only if `keep_device`. */
static void lazy_open(struct lazy_range* g, int keep_device){
    struct shim_state* s = g->s;
    const int was = g->prot;
    g->prot = LAZY_OFF;
    mprotect(g->lo, (size_t)(g->hi-g->lo), PROT_READ|PROT_WRITE);
    if (was==LAZY_NONE && s->host_stale){
        const uint64_t n = rebcu_N(s->h);
        if (rebcu_download(s->h, (rebcu_particle*)g->lo, n)==0){ s->host_stale = 0; s->uploaded_N = n; }
    }
    s->lazy = 0;
    if (!keep_device) s->device_valid = 0;
}

static void lazy_handler(int sig, siginfo_t* si, void* ctx){
    struct lazy_range* g = find_addr(si->si_addr);
    if (g){
        struct shim_state* s = g->s;
        if (g->prot==LAZY_NONE){
            /* first touch: fetch the particles, allow reads */
            mprotect(g->lo, (size_t)(g->hi-g->lo), PROT_READ|PROT_WRITE);
            if (s->host_stale){
                const uint64_t n = rebcu_N(s->h);
                if (rebcu_download(s->h, (rebcu_particle*)g->lo, n)==0){ s->host_stale = 0; s->uploaded_N = n; }
            }
            s->lazy_faults++;
            g->prot = LAZY_READ;
            mprotect(g->lo, (size_t)(g->hi-g->lo), PROT_READ);
            return;                                   /* the faulting instruction is retried */
        }
        if (g->prot==LAZY_READ){
            /* a write: the host copy becomes the truth */
            g->prot = LAZY_OFF;
            mprotect(g->lo, (size_t)(g->hi-g->lo), PROT_READ|PROT_WRITE);
            s->lazy = 0;
            s->device_valid = 0;
            return;
        }
    }
    /* not ours: hand the fault to whoever was there before */
    if (old_action.sa_flags & SA_SIGINFO){ if (old_action.sa_sigaction){ old_action.sa_sigaction(sig, si, ctx); return; } }
    else if (old_action.sa_handler==SIG_IGN) return;
    else if (old_action.sa_handler!=SIG_DFL && old_action.sa_handler){ old_action.sa_handler(sig); return; }
    signal(SIGSEGV, SIG_DFL);                         /* default action on the retried instruction */
}

/* Can r->particles be protected?  It must own whole pages (see reb_b200_particles_realloc) and be worth it. */
int shim_lazy_possible(const struct reb_simulation* r){
    if (!lazy_enabled() || !r->particles) return 0;
    const size_t bytes = r->N_allocated*sizeof(struct reb_particle);
    if (bytes < LAZY_MIN_BYTES) return 0;
    if (((uintptr_t)r->particles & (page-1)) != 0) return 0;
    const size_t rounded = (bytes + page-1) & ~(page-1);
    return malloc_usable_size(r->particles) >= rounded;
}

/* After a device step under a heartbeat: hide the (now stale) host copy. */
void shim_lazy_protect(struct reb_simulation* r, struct shim_state* s){
    if (!shim_lazy_possible(r) || s->pinned_ptr) return;
    pthread_mutex_lock(&lazy_lock);
    if (!handler_installed){
        struct sigaction sa;
        memset(&sa, 0, sizeof(sa));
        sa.sa_sigaction = lazy_handler;
        sa.sa_flags = SA_SIGINFO | SA_NODEFER;
        sigemptyset(&sa.sa_mask);
        if (sigaction(SIGSEGV, &sa, &old_action)==0) handler_installed = 1;
    }
    struct lazy_range* g = NULL;
    if (handler_installed){
        for (int i=0;i<n_ranges && !g;i++) if (ranges[i].s==s) g = &ranges[i];
        for (int i=0;i<n_ranges && !g;i++) if (ranges[i].s==NULL) g = &ranges[i];
        if (!g && n_ranges<LAZY_MAX) g = &ranges[n_ranges++];
    }
    pthread_mutex_unlock(&lazy_lock);
    if (!g) return;
    const size_t bytes = r->N_allocated*sizeof(struct reb_particle);
    const size_t rounded = (bytes + page-1) & ~(page-1);
    if (g->prot!=LAZY_OFF && g->lo!=(char*)r->particles) mprotect(g->lo, (size_t)(g->hi-g->lo), PROT_READ|PROT_WRITE);
    g->s = s; g->lo = (char*)r->particles; g->hi = g->lo + rounded;
    if (mprotect(g->lo, rounded, PROT_NONE)==0){ g->prot = LAZY_NONE; s->lazy = 1; }
    else { g->prot = LAZY_OFF; s->lazy = 0; }
}

/* Before the shim (or libc on its behalf) touches r->particles: lift the protection, host copy current. */
void shim_lazy_open(struct shim_state* s, int keep_device){
    if (!s || !s->lazy) return;
    for (int i=0;i<n_ranges;i++) if (ranges[i].s==s && ranges[i].prot!=LAZY_OFF){ lazy_open(&ranges[i], keep_device); return; }
    s->lazy = 0;
}

void shim_lazy_forget(struct shim_state* s){
    for (int i=0;i<n_ranges;i++) if (ranges[i].s==s){
        if (ranges[i].prot!=LAZY_OFF) mprotect(ranges[i].lo, (size_t)(ranges[i].hi-ranges[i].lo), PROT_READ|PROT_WRITE);
        ranges[i].prot = LAZY_OFF; ranges[i].s = NULL; ranges[i].lo = ranges[i].hi = NULL;
    }
}

#pragma GCC visibility push(default)
/* realloc / free of the reference's particle.c and binarydata.c (-D renames): page-owning blocks for large arrays, and
 * no libc access to a protected block. */
void* reb_b200_particles_realloc(void* ptr, size_t size){
    lazy_enabled();
    if (ptr){
        struct lazy_range* g = find_addr(ptr);
        if (g && g->lo==(char*)ptr) lazy_open(g, 1);           /* the copy below reads the block */
    }
    if (size < LAZY_MIN_BYTES || !lazy_enabled()) return realloc(ptr, size);
    const size_t rounded = (size + page-1) & ~(page-1);
    void* p = NULL;
    if (posix_memalign(&p, page, rounded)) return realloc(ptr, size);
    if (ptr){
        const size_t old = malloc_usable_size(ptr);
        memcpy(p, ptr, old < size ? old : size);
        free(ptr);
    }
    return p;
}

void reb_b200_particles_free(void* ptr){
    if (ptr){
        struct lazy_range* g = find_addr(ptr);
        if (g && g->lo==(char*)ptr){
            /* the block is being discarded: no download, just make it ordinary memory again */
            mprotect(g->lo, (size_t)(g->hi-g->lo), PROT_READ|PROT_WRITE);
            g->prot = LAZY_OFF; if (g->s){ g->s->lazy = 0; g->s->host_stale = 0; g->s->device_valid = 0; }
        }
    }
    free(ptr);
}
#pragma GCC visibility pop
