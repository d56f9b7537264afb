// launch.h -- plain-old-data descriptors shared by the host context and the
// kernel translation units, and the launcher entry points each instantiation
// unit (sweep_inst.cu compiled once per space dimension and base method)
// exports.
#pragma once
#include <cuda_runtime.h>

#include "numerics.cuh"

namespace fvs {

// One box (patch) of a level as the sweep kernels see it.  All indices are
// relative to the lower corner of the state array (which includes ghosts).
struct BoxDesc {
  const double* src; // state read by the kernel, ncomp components
  double* dst;       // state written by the kernel (fused sweep), or == src
  double* flux;      // flux array of the sweep direction (unfused kernels)
  int n[3];          // extents of the state array
  int fn[3];         // extents of the flux array
  int foff[3];       // flux index = cell index - foff (per dimension)
  long long sc;      // component stride of the state array
  long long fsc;     // component stride of the flux array
  int ulo, uhi;      // cells updated along the sweep direction (inclusive)
  int flo, fhi;      // faces computed along the sweep direction (face f lies between cells f-1 and f)
  const unsigned char* dtmask; // n[dir] bytes: dtmask[i] != 0 iff face (i | i+1) enters the stable dt
  // embedded boundary (levels with fvs_level_set_eb only; null / 0 otherwise)
  const double* alpha; // volume fractions, extents n[]
  const double* beta;  // face fractions of the sweep direction, extents n[] + 1 along dir
  int fab_type;        // -1 covered, 0 regular, 1 single valued
};

// A step of a fused sweep whose branch-free fast paths left their domain: cells lo..hi of the
// grid line that starts at element offset `line` of box `box` are recomputed by the repair
// pass that follows the sweep (sweep_kernels.cuh, "Mark and repair").
struct RepairEntry {
  long long line;
  int box;
  int lo, hi;
};

struct SweepArgs {
  EqConst eq;
  double lambda;      // dt / dx
  double lambda_half; // (0.5 * dt) / dx
  const BoxDesc* boxes; // device array
  int nboxes;
  int max_n[3];       // max extents over the boxes (grid sizing)
  int seg_len;        // output cells per segment along the sweep direction
  unsigned long long* max_signal; // device: bit pattern of the max signal speed (non-negative double)
  int* fail_flag;     // device: set to 1 if a Newton iteration failed
  // lazy stable dt (numerics.cuh "Lazy stable dt"): a value the caller believes
  // to be <= the max signal speed of this sweep; faces whose cheap bound stays
  // below it are not evaluated.  The caller MUST verify result >= sig_hint
  // afterwards and repeat with 0 otherwise.  0 = no hint.
  double sig_hint;
  // mark and repair: list of the steps to recompute, its length ([0]; [1] counts the blocks of
  // the repair pass that are done) and its capacity -- a longer list means "every cell"
  RepairEntry* repair_list;
  unsigned* repair_count;
  unsigned repair_cap;
};

struct MethodKey {
  int rank, base, recon, limiter;
};

// Each returns the number of kernels launched, or -1 if the combination is
// not compiled into this unit.
using SweepLauncher = int (*)(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);

// fused: stable dt + numeric flux + conservative update + complete, src -> dst
int launch_fused_sweep(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);
// the same with the tolerance arithmetic (fvs_set_arithmetic); -1 for the exact Riemann solver
int launch_fused_sweep_tolerance(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);
// unfused policies (operate on src in place, fluxes through BoxDesc::flux)
int launch_stable_dt(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);
int launch_numeric_fluxes(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);
int launch_update_conservatively(int rank, int dir, const SweepArgs&, cudaStream_t);
int launch_complete_from_cons(int rank, const SweepArgs&, cudaStream_t);
// stable dt of all directions in one pass (masks: one per direction)
struct DtAllArgs {
  EqConst eq;
  const BoxDesc* boxes[3]; // per direction (dtmask differs)
  int nboxes;
  int max_n[3];
  unsigned long long* max_signal; // [3] one accumulator per direction
  int* fail_flag;
  double sig_hint[3]; // per direction, see SweepArgs::sig_hint
};
int launch_stable_dt_all(const MethodKey&, const DtAllArgs&, cudaStream_t);
// the same with the tolerance arithmetic (regular levels of a context that opted in)
int launch_stable_dt_all_tolerance(const MethodKey&, const DtAllArgs&, cudaStream_t);

// ---- cut cells (embedded boundary) ------------------------------------------
// masked regular fluxes (KbnCutCellMethod::ComputeRegularFluxes) and the
// cut-cell adapter's stable dt, per method instantiation unit
int launch_cc_numeric_fluxes(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);
int launch_cc_stable_dt(const MethodKey&, int dir, const SweepArgs&, cudaStream_t);

// One box of an EB level as the method-independent cut-cell kernels see it.
struct EbDesc {
  int fab_type;
  int n[3];      // state array extents
  long long sc;  // state component stride
  int fn[3], foff[3];
  long long fsc; // flux arrays of the sweep direction
  int bn[3];
  long long bsc; // face-fraction arrays of the sweep direction (extents n + e_dir)
  int ulo, uhi, flo, fhi;
  double* state;
  const double* alpha;
  const double* beta;
  const double* bcent; // D components, stride sc
  const double* bnorm; // D components, stride sc
  const double* us;    // 2 components (absolute, relative), stride bsc
  const double* sl;
  const double* sr;
  double* reference;     // Complete components, stride sc
  double* boundary_flux; // Conservative components, stride sc
  double* flux;          // regular ("unshielded") fluxes, stride fsc
  double* stab;
  double* shl;
  double* shr;
  const long long* cut_cells; // compacted list: state-array offsets of the cut cells
  int n_cut;
  const unsigned char* complete_mode; // per cell: 0 skip, 1 plain, 2 masked by alpha
};
struct EbArgs {
  EqConst eq;
  const EbDesc* boxes;
  int nboxes;
  int rank, dir;
  int riemann; // 0 HLL + Einfeldt, 1 exact
  double lambda;
  long long max_cells; // grid sizing
  int max_cut;
  int* fail_flag;
};
int launch_cc_reference_states(const EbArgs&, cudaStream_t);
int launch_cc_boundary_fluxes(const EbArgs&, cudaStream_t);
int launch_cc_stabilised_fluxes(const EbArgs&, cudaStream_t);
int launch_cc_update(const EbArgs&, cudaStream_t);
int launch_cc_complete(const EbArgs&, cudaStream_t);
// FillCutCellData for one direction of one box
int launch_fill_cutcell_data(const double* beta, double* us, double* sl, double* sr, double* ds,
                             const int bn[3], int dir, cudaStream_t);
int launch_cut_cell_flags(const double* alpha, const double* const beta[3], const int n[3],
                          int rank, unsigned char* flags, cudaStream_t);
// tagging
int launch_tag_gradient(const double* state, long long sc, const int n[3], const double* alpha,
                        int rank, const int lo[3], const int len[3], int ncond, const int* comps,
                        const double* tols, char* tags, cudaStream_t);
int launch_tag_cutcells(const double* alpha, const int n[3], const int lo[3], const int len[3],
                        char* tags, cudaStream_t);
int launch_tag_buffer(char* tags, char* scratch, int rank, const int n[3], int width,
                      cudaStream_t);

// ---- coarse-fine operations (two levels on one rank) -------------------------
// All index triples are GLOBAL cell / face indices of the level they refer to;
// an array is addressed through its lower corner `*_lo` and extents `*_n`.
struct ArrayRef {
  double* p;
  int lo[3], n[3];
  long long sc; // component stride
};
struct AvgDownDesc { // coarse region <- mean of the fine cells, conservative comps
  ArrayRef fine, crse;
  int lo[3], len[3]; // coarse region
  // volume fractions on the fine array's box: amrex::EB_average_down of a
  // single-valued fab (sum(fine * vfrac) / sum(vfrac)); nullptr = plain mean
  const double* vfrac;
};
struct InterpDesc { // fine ghost region <- coarse parent (pc_interp)
  ArrayRef fine, crse;
  int lo[3], len[3];   // fine region
  int vlo[3], vhi[3];  // the fine box's own valid cells (skipped)
};
struct RegisterDesc { // one side of one fine box, coarse faces of direction dir
  ArrayRef reg;                // plane of ncons components
  const unsigned char* open;   // Reflux mask
  int side;
  ArrayRef other;              // FineAdd: fine flux array; CrseInit: coarse flux; Reflux: coarse state
  int box_lo[3], box_hi[3];    // CrseInit / Reflux: the coarse box's valid cells
  int lo[3], len[3];           // faces of the register handled by this descriptor
};
int launch_average_down(const AvgDownDesc* d, int n, int ncons, const int ratio[3], long long max_cells,
                        cudaStream_t);
int launch_pc_interp(const InterpDesc* d, int n, int ncomp, const int ratio[3], long long max_cells,
                     cudaStream_t);
// op: 0 FineAdd, 1 CrseInit, 2 Reflux
int launch_flux_register(const RegisterDesc* d, int n, int op, int ncons, int dir, const int ratio[3],
                         double mult, double vol, long long max_faces, cudaStream_t);

// ---- ghost cells -----------------------------------------------------------
struct CopyDesc {
  const double* src;
  double* dst;
  int sn[3];
  int dn[3]; // extents of source / destination arrays
  long long ssc, dsc;
  int slo[3]; // first source cell (array-relative)
  int dlo[3]; // first destination cell (array-relative)
  int len[3]; // region extents
};
int launch_copy_regions(const CopyDesc* regions, int nregions, int ncomp, long long max_cells,
                        cudaStream_t);
// destinations in a peer's memory: 16-byte pieces, four in flight per thread
int launch_copy_regions_wide(const CopyDesc* regions, int nregions, int ncomp, long long max_cells,
                             cudaStream_t);
// pack / unpack of regions to / from a contiguous buffer (multi-GPU halo)
struct PackDesc {
  double* array; // state array
  int n[3];
  long long sc;
  int lo[3], len[3];
  long long offset; // first element in the packed buffer
};
int launch_pack(const PackDesc* regions, int nregions, int ncomp, double* buffer,
                long long max_cells, int unpack, cudaStream_t);

struct BcDesc {
  double* array;
  int n[3];
  long long sc;
  int lo[3], len[3]; // fill region (array-relative)
  int src_plane;     // transmissive: array-relative index of the boundary cell along dir
  int mirror_sum;    // reflective: src = mirror_sum - idx along dir
};
int launch_boundary(const BcDesc* regions, int nregions, int ncomp, int dir, int kind,
                    long long max_cells, cudaStream_t);

} // namespace fvs
