/* chimp_b200.h -- C-ABI of the B200 lattice-Boltzmann engine.
 *
 * The reference (eje74/BADChIMP-cpp) has no FFI: its "API" is the header set pulled in by
 * src/LBSOLVER.h, consumed by hand-written mains.  Each entry point below names the
 * reference interface it stands in for.  All pointers are plain host pointers unless a
 * name says `_dev`; arrays are borrowed for the duration of the call only; every function
 * returns 0 on success and non-zero on error (message via chimp_last_error()).  The
 * reference convention (print + exit(1), e.g. LBvtk.h:230-233) is left to the host wrapper.
 *
 * One chimp_lattice per (rank, GPU).  Not thread-safe.  Step calls are asynchronous on the
 * context's stream; download / moment / reduction calls synchronise.
 *
 * Environment switches, read once when a lattice is created:
 *   CHIMP_PEER_TIMEOUT_MS  bound of every device-side wait for a neighbour rank (default 20000); a wait that gives up
 *                          raises an error that the next synchronising call returns
 *   CHIMP_PEER_FUSED=0     peer halos through separate push launches instead of inside the step kernel
 *   CHIMP_TRACE=1          per-rank phase timings of N-rank stepping on stderr
 *   CHIMP_ATTR_PACKED=0    one_phase step kernel reads its per-node attributes from four arrays instead of one packed word
 *   CHIMP_PHI_DERIVED=0    two-phase collide pass reads the full phi table (4 nQ bytes per node) instead of its derived form
 */
#ifndef CHIMP_B200_H
#define CHIMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct chimp_lattice chimp_lattice;

/* lattice ids: D2Q9 (LBd2q9.h:12), D3Q19 (LBd3q19.h:12), D3Q27 (new, same contract) */
enum { CHIMP_D2Q9 = 0, CHIMP_D3Q19 = 1, CHIMP_D3Q27 = 2 };
/* collision operators: calcOmegaBGK (LBcollision.h:27), calcOmegaBGKTRT (LBcollision.h:50) */
enum { CHIMP_BGK = 0, CHIMP_TRT = 1 };
/* index forms of the streaming step */
enum { CHIMP_INDEX_TABLE = 0, CHIMP_INDEX_COMPACT = 1 };
/* link boundary kinds of std_one_phase/main.cpp:138-203 */
enum { CHIMP_LINK_SOLID = 0, CHIMP_LINK_PRESSURE = 1, CHIMP_LINK_FLUID_SWAP = 2 };

const char *chimp_last_error(void);
int chimp_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
long long chimp_launch_count(void);
/* lattice constants, as the reference structs expose them (LBd3q19.h:14-42) */
int chimp_lattice_nq(int lattice);
int chimp_lattice_nd(int lattice);
int chimp_lattice_c(int lattice, int q, int d);
double chimp_lattice_w(int lattice, int q);
int chimp_lattice_reverse(int lattice, int q);

/* ---- construction: mirrors Grid<L>(vtk) + findBulkNodes (LBgrid.h:127-148, LBgeometry.h:11-21).
 * neigh is Grid::neigList_ in reference layout [n_nodes * nQ] (LBgrid.h:187-198), bulk is the
 * ascending list of own fluid labels, n_fields is LbField's nFields (1, or 2 for twophase).
 * device < 0 uses the current CUDA device. */
int chimp_create(chimp_lattice **out, int lattice, int n_nodes, const int32_t *neigh, int n_bulk,
                 const int32_t *bulk, int n_fields, int device);

/* HalfWayBounceBack<L>(bndNodes, nodes, grid) + apply (LBhalfwaybb.h:25-63): per boundary node
 * the beta (unknown) directions followed by gamma and delta pair directions, nDirPairs entries
 * per node as in BoundaryHalwWayHelper::linkList_ (LBhalfwayhelperclass.h:163-212). */
int chimp_add_halfway_bb(chimp_lattice *, int n_bnd, const int32_t *nodes, const int32_t *n_beta,
                         const int32_t *n_gamma, const int32_t *n_delta, const int32_t *links);

/* link lists {nodeFluid, qUnknown, nodeWall, qKnown} of std_one_phase/main.cpp:27-126, applied
 * after streaming in the order added (main.cpp:591-597). */
int chimp_add_links(chimp_lattice *, int kind, int n_links, const int32_t *links4);

/* PressureBnd<DXQY>::apply / InletOutlet<DXQY>::apply (LBpressurebnd.h:10-88): after the boundary phase of every step the
 * population f(q_k, node_k) holds value_k.  node_q = n_links pairs (destination node label, direction); for the
 * reference classes the destination is grid.neighbor(beta, bndNode) with value w[beta] * rho(bndNode)  (PressureBnd) or
 * rho * w[beta] * (1 + c2Inv cu + c4Inv0_5 (cu^2 - c2 u^2))  (InletOutlet); the host mirror host/chimp/LBpressurebnd.h
 * forms exactly those products.  The values are captured here (a prescribed density / velocity; the reference has no
 * caller that varies them).  Call in the order the main applies its boundaries, before chimp_finalize; one-field
 * lattices only.  Downloads return the constants at those places, like the reference's field after apply(). */
int chimp_add_constant_links(chimp_lattice *, int n_links, const int32_t *node_q, const double *values);

/* one MonLatMpi (LBmonlatmpi.h:72-98): ghost exchange lists towards one neighbour rank, in
 * the reference's list order.  Neighbours must be added in ascending rank order (LBvtk.h:534-544). */
int chimp_add_neighbor(chimp_lattice *, int neig_rank, int n_send, const int32_t *nodes_to_send,
                       const int32_t *n_dir_send, const int32_t *dir_list_send, int n_recv,
                       const int32_t *nodes_received, const int32_t *n_dir_recv, const int32_t *dir_list_recv);

/* scalar-field support rows for twophase: solid boundary nodes (findSolidBndNodes,
 * LBgeometry.h:37-45) carry a constant colour value derived from the wettability densities. */
int chimp_set_solid_boundary(chimp_lattice *, int n_solid, const int32_t *solid_nodes);

/* host half of the build: symbolic replay of one reference iteration -> pull table and halo
 * lists.  Pure host code (no CUDA call); chimp_finalize runs it if it has not been run. */
int chimp_build_host(chimp_lattice *, int boundary_first);
/* host tables for inspection: info6 = {n_own, n_pad, n_halo, plane_stride, n_boundary, 0};
 * table int32 [nQ][n_own] (source slot, -1 = reversed own slot), labels [n_own], pmask [n_own];
 * halo lists of neighbour k as slot offsets q*plane_stride + slot. */
int chimp_host_table_info(chimp_lattice *, long long *info6);
int chimp_host_table(chimp_lattice *, int32_t *table, int32_t *labels, uint32_t *pmask);
int chimp_host_halo_lists(chimp_lattice *, int k, long long *send_src, long long *recv_dst);
/* the constant links (chimp_add_constant_links) in registration order: the slot offset q*plane_stride + slot each one
 * occupies and its value; returns their number (either pointer may be null), -1 on error */
int chimp_host_constant_links(chimp_lattice *, long long *dst, double *values);
/* the same for the scalar (phi) halo of a two-field lattice: phi slots sent / ghost phi slots received */
int chimp_host_scalar_halo_lists(chimp_lattice *, int k, long long *send_src, long long *recv_dst);

/* builds the device tables; index_form is CHIMP_INDEX_TABLE or CHIMP_INDEX_COMPACT.
 * boundary_first != 0 orders halo-coupled nodes first so that their step can overlap. */
int chimp_finalize(chimp_lattice *, int index_form, int boundary_first);

/* alternative construction from device-resident pull tables (structured geometry ingest,
 * replaces vtklb.py -> ASCII -> LBvtk for large cases): table_dev is int32 [nQ][n_pad]
 * (source slot, -1 = reversed own slot), label_dev int32 [n_pad] reference labels. */
int chimp_create_from_device_table(chimp_lattice **out, int lattice, int n_bulk, int n_pad, int n_halo,
                                   const int32_t *table_dev, const int32_t *label_dev, int n_fields,
                                   int index_form, int device);

/* Structured geometry ingest for HOST code (C / C++ mains without torch): one rank's lattice straight from a raw
 * voxel array -- replaces vtklb.py -> ASCII file -> LBvtk (LBvtk.h:221-262), whose int offsets stop at 2 GiB of text
 * (LBvtk.h:194-201), so a main reaches 512^3 and beyond.  voxels is uint8 [nx][ny][nz] in C-order (a 2-D lattice takes
 * [nx][ny], nz = 1), 0 = solid; periodic_mask bit 0/1/2 = x/y/z periodic, other axes are closed.  Own fluid nodes
 * get the reference's labels 1..N in C-order (vtklb.py:92-94), so chimp_upload_lbfield / chimp_download_* work on
 * arrays of N + 1 rows; std_case semantics: half-way bounce back on every link into a solid (LBhalfwaybb.h:37-63).
 * chimp_voxel_table_host is the host half alone (no CUDA call): sizes, and -- when the pointers are not NULL -- the
 * pull table int32 [nQ][n_pad] and labels [n_pad].  chimp_set_phi_table_from_voxels adds the colour-gradient tables
 * of a two-field lattice (wall colour wall_phi[cell] at the solid cells next to fluid, main_TWOPHASE.cpp:280-284). */
int chimp_voxel_table_host(int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask, int *n_own, int *n_pad,
                           int32_t *table, int32_t *labels);
int chimp_create_from_voxels(chimp_lattice **out, int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask,
                             int n_fields, int index_form, int device);
int chimp_voxel_phi_table_host(int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask, const double *wall_phi,
                               int *n_extra, int32_t *ptable, double *phi_extra);
int chimp_set_phi_table_from_voxels(chimp_lattice *, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask, const double *wall_phi);

/* One z-slab of an N-rank decomposition for host code: voxels_ext is uint8 [nx][ny][nz_own + 2] -- the rank's own
 * layers plus one layer of the rank below (z index 0) and above (z index nz_own + 1); x and y are periodic, the ranks
 * form a ring in z (rank map of relperm_input.py:16-35 with nproc = [1, 1, N]).  Own fluid cells get the reference's
 * per-rank labels 1..N; the two halo-coupled layers occupy the leading device slots, the others follow layer by layer.
 * The lattice comes with its two halo faces registered (face 0 towards rank_down, face 1 towards rank_up) and the
 * boundary count set; what remains is the transport: chimp_connect_peer per face with the neighbour's handles /
 * pointers and ITS receive list (chimp_halo_face_recv_list of the face that looks at me), or the exchange callback.
 * chimp_slab_tables_host is the host half alone (no CUDA call): info8 = {n_own, n_pad, n_halo, n_boundary, n_send_down,
 * n_recv_down, n_send_up, n_recv_up}; the other pointers may be NULL. */
int chimp_slab_tables_host(int lattice, int nx, int ny, int nz_own, const uint8_t *voxels_ext, long long *info8, int32_t *table,
                           int32_t *labels, long long *send_down, long long *recv_down, long long *send_up, long long *recv_up);
int chimp_create_slab_from_voxels(chimp_lattice **out, int lattice, int nx, int ny, int nz_own, const uint8_t *voxels_ext, int n_fields,
                                  int index_form, int device, int rank_down, int rank_up);
long long chimp_halo_face_recv_count(chimp_lattice *, int k);
int chimp_halo_face_recv_list(chimp_lattice *, int k, long long *recv_dst);

void chimp_destroy(chimp_lattice *);

/* ---- state transfer in reference layout and labels.
 * f_aos is LbField::data_ (LBfield.h:300): [(n_fields*nQ)*node + nQ*field + q], n_nodes rows;
 * only own bulk rows are read / written. */
int chimp_upload_lbfield(chimp_lattice *, const double *f_aos);
int chimp_download_lbfield(chimp_lattice *, double *f_aos);
/* ScalarField rho (LBfield.h:94: [nFields*node + field]) and VectorField vel (LBfield.h:190) as
 * stored by the last step that was run with moments enabled. */
int chimp_download_rho(chimp_lattice *, double *rho_sca, int n_fields_host);
int chimp_download_vel(chimp_lattice *, double *vel_vec);

/* ---- per-node attributes of std_one_phase (main.cpp:253-330), indexed by reference label */
int chimp_set_one_phase_attributes(chimp_lattice *, const double *force_on, const int32_t *interior_label,
                                   const double *add_mass_source, int n_labels, const double *scale_per_label,
                                   double rho_w);

/* ---- stepping.  n_steps full iterations (collide, stream, exchange, boundary); rho / vel
 * are stored on the last step only (they are the moments the reference holds after the same
 * number of iterations).
 * chimp_step_single: std_case/main.cpp:109-146 (BGK or TRT + Guo force + bounce back);
 * with one-phase attributes set it runs std_one_phase/main.cpp:513-597. */
typedef struct {
    int collision;          /* CHIMP_BGK / CHIMP_TRT */
    double tau;             /* BGK */
    double tau_sym, tau_anti; /* TRT */
    double force[3];        /* bodyForce(0,0) */
} chimp_single_params;
int chimp_step_single(chimp_lattice *, const chimp_single_params *, int n_steps);
/* the two halves of one iteration, for hosts that move the halos themselves between them:
 * begin = collide/stream of all own nodes + packing of outgoing populations (LBmonlatmpi.h:244-252),
 * end   = unpacking of incoming ones (LBmonlatmpi.h:259-267) + buffer swap (LBfield.h:359) */
int chimp_step_begin(chimp_lattice *, const chimp_single_params *, int store_moments);
int chimp_step_end(chimp_lattice *);
/* massChange[label] of the last step (std_one_phase/main.cpp:528) */
int chimp_download_mass_change(chimp_lattice *, double *mass_per_label);
/* same, bracketed by CUDA events on the engine's stream; *ms = device time of the n_steps */
int chimp_step_timed(chimp_lattice *, const chimp_single_params *, int n_steps, double *ms);
/* state f_q(n) = w_q * rho for every own node (std_case/main.cpp:92-96 with constant rho);
 * works for lattices created from device tables, where no reference-layout upload exists */
int chimp_init_uniform(chimp_lattice *, double rho);
/* structured-ingest lattices: state f_{s,q}(i) = w_q * rho_s(i) (initiateLbField with u = 0,
 * LBinitiatefield.h:33-57) from a device array rho [n_fields][n_own] in device node order */
int chimp_init_equilibrium_dev(chimp_lattice *, const double *rho_dev);
/* structured-ingest two-field lattices: the phi slot of neighbor(q, n) for every own node
 * (int32 [nQ][n_pad]: own slot, n_pad + k for extra slot k, n_pad + n_extra for "always 0") and the constant
 * colour of the n_extra wall slots (main_TWOPHASE.cpp:280-284); replaces chimp_set_solid_boundary +
 * chimp_set_twophase_density of the reference-table path */
int chimp_set_phi_table_dev(chimp_lattice *, const int32_t *ptable_dev, int n_extra, const double *phi_extra_dev);
/* rho [n_own] and vel [nD][n_own] of the last step, in device node order */
int chimp_download_moments_device_order(chimp_lattice *, double *rho, double *vel);

/* twophase/main_TWOPHASE.cpp:236-392 (colour gradient, 2 fields, flux-controlled force) */
typedef struct {
    double tau0, tau1, sigma, beta, momx;
    double force[3];        /* y,z components; x is overwritten by the flux controller */
    long long n_fluid_global; /* numNodesGlobal (main_TWOPHASE.cpp:196-198) */
} chimp_twophase_params;
int chimp_set_twophase_density(chimp_lattice *, const double *rho_sca2); /* ScalarField rho(2,size) incl. wall rows */
int chimp_step_twophase(chimp_lattice *, const chimp_twophase_params *, int n_steps);
/* same, bracketed by CUDA events on the engine's stream; *ms = device time of the n_steps */
int chimp_step_twophase_timed(chimp_lattice *, const chimp_twophase_params *, int n_steps, double *ms);
int chimp_download_phase_field(chimp_lattice *, double *cg_sca);
double chimp_last_flux_force(chimp_lattice *);

/* ---- on-device reductions for callers (the step either side of the path) ----
 * chimp_flux_force: calcFluxForceCartDir (LBglobalforcing.h:8-33): 2*(fixed_flux - sum_n qSumC(f(field,n))[cart_dir] / N)
 * over the own nodes of all ranks (fixed-shape tree sum on the device, all-reduce callback across ranks).
 * chimp_node_list_flux: out[bin[k]] += vel(0, component, nodes[k]) * rho(field, nodes[k]) in list order -- the
 * mass flux through the pressure-boundary nodes (std_one_phase/main.cpp:607-619, bin = fluid phase, component = 2);
 * uses the moments of the last step of the previous chimp_step_* call; only the products of the listed nodes
 * cross the bus.  nodes are reference labels. */
int chimp_flux_force(chimp_lattice *, int field_no, int cart_dir, double fixed_flux, long long n_nodes_global, double *force_out);
/* calcCapNumbForceCartDir (LBglobalforcing.h:35-98) on a two-field lattice: uniform force that fixes the capillary number,
 * 2*(sigma_cap_numb - (<phi0 m> nu0 + <phi1 m> nu1)) / (<phi0> nu0 + <phi1> nu1), m = qSumC(f(0,n))[cart_dir],
 * phi_s = rho_s/(rho_0+rho_1) from the moments of the last step, <.> = sum over all ranks' own nodes / n_nodes_global. */
int chimp_capillary_force(chimp_lattice *, int cart_dir, double sigma_cap_numb, double nu0, double nu1, long long n_nodes_global,
                          double *force_out);
int chimp_node_list_flux(chimp_lattice *, int n_list, const int32_t *nodes, const int32_t *bin, int n_bins, int field_no,
                         int component, double *out);

/* ---- halo plumbing for N ranks (replaces MonLatMpi::communicateLbField, LBmonlatmpi.h:236-297).
 * The engine packs outgoing populations into per-neighbour device buffers and unpacks
 * incoming ones; the transport (NCCL send/recv or peer stores) is supplied by the host. */
int chimp_num_neighbors(chimp_lattice *);
int chimp_neighbor_info(chimp_lattice *, int k, int *neig_rank, long long *send_count, long long *recv_count);
/* device pointers of the packed send / recv buffers of neighbour k (doubles, per field contiguous) */
void *chimp_send_buffer_dev(chimp_lattice *, int k);
void *chimp_recv_buffer_dev(chimp_lattice *, int k);
/* let the engine pack into / unpack from caller-owned device buffers (e.g. torch tensors used with NCCL) */
int chimp_set_halo_buffers(chimp_lattice *, int k, void *send_dev, void *recv_dev);
void *chimp_halo_stream(chimp_lattice *);
/* structured-ingest lattices (chimp_create_from_device_table): one halo face towards neig_rank.
 * send_src / recv_dst are slot offsets q*plane_stride + slot inside one field (host arrays);
 * their order defines the packed message.  The same rank may own two faces (2-rank ring). */
int chimp_add_halo_face(chimp_lattice *, int neig_rank, long long n_send, const long long *send_src,
                        long long n_recv, const long long *recv_dst);
/* structured-ingest two-field lattices: the scalar (phi) halo of neighbour / face k (communciateScalarField,
 * LBbndmpi.h:141-156 -> LBmonlatmpi.h:181-205): phi slots of the own nodes whose colour the neighbour needs, in
 * message order, and the ghost phi slots (>= n_pad, as numbered in chimp_set_phi_table_dev) the incoming values
 * go to.  send_dev / recv_dev: caller-owned device buffers for the transport (both or neither). */
int chimp_add_scalar_halo_face(chimp_lattice *, int k, long long n_send, const long long *send_src, long long n_recv,
                               const long long *recv_dst, void *send_dev, void *recv_dev);
/* Peer halos: instead of pack -> host transport -> unpack, the engine stores the outgoing populations
 * directly into the neighbour GPU's halo-in slots over NVLink and publishes an arrival counter there
 * (fused into the halo-coupled part of the step; replaces MPI_Send/MPI_Recv of LBmonlatmpi.h:253-257).
 * chimp_ipc_handles: 3 x 64-byte CUDA IPC handles (population buffer A, buffer B, arrival flags) for the
 * other process; chimp_connect_peer: what neighbour/face k of THIS lattice needs to know about its peer --
 * the peer's handles (or, when both lattices live in one process, its raw pointers from
 * chimp_local_pointers), the peer's field stride nQ*plane_stride, the index of the peer's face that I feed,
 * and for each entry of my send list the slot offset q*peer_plane_stride + slot it lands in (the peer's
 * receive list).  Once every neighbour is connected, stepping uses the peer path.  All ranks must step
 * the same number of times. */
int chimp_ipc_handles(chimp_lattice *, unsigned char *out192);
int chimp_local_pointers(chimp_lattice *, void **out3);
int chimp_connect_peer(chimp_lattice *, int k, const unsigned char *peer_handles192, int same_process,
                       void *const *peer_ptrs3, long long peer_field_stride, int peer_face, long long n_dst,
                       const long long *peer_dst);
/* 0: no peer halos; 1: peer halos through separate push launches; 2: fused into the step kernel (one launch per step).
 * `why` receives the reason when the fused form could not be used. */
int chimp_peer_mode(chimp_lattice *, char *why, int why_len);
/* Two-phase lattices over peer memory: additionally the scalar halo of phi is stored into the neighbours' ghost slots
 * and the momentum sum of the flux controller (MPI_Allreduce, main_TWOPHASE.cpp:299) travels through a mailbox every rank
 * exports: chimp_ipc_handles_twophase gives 2 x 64-byte handles (phi array, mailbox); chimp_connect_peer_scalar (after
 * chimp_connect_peer of the same face) takes the neighbour's phi handle (or pointer, same process) and, per entry of my
 * scalar send list, the ghost phi slot of the neighbour it lands in; chimp_connect_world takes the mailbox handles
 * (world x 64 bytes, or pointers) of all ranks in rank order.  The sums are added in rank order on every rank, so all
 * ranks obtain identical bits.  Afterwards chimp_step_twophase needs no callbacks. */
int chimp_ipc_handles_twophase(chimp_lattice *, unsigned char *out128);
int chimp_local_pointers_twophase(chimp_lattice *, void **out2);
int chimp_connect_peer_scalar(chimp_lattice *, int k, const unsigned char *peer_phi_handle64, int same_process, void *peer_phi_ptr,
                              long long n_dst, const long long *peer_phi_dst);
int chimp_connect_world(chimp_lattice *, int rank, int world, const unsigned char *mail_handles, int same_process, void *const *mail_ptrs);
/* the first n_boundary slots hold the halo-coupled nodes: they are stepped and packed first */
int chimp_set_boundary_count(chimp_lattice *, int n_boundary);
typedef int (*chimp_exchange_fn)(void *user, void *stream);
/* called once per step after the boundary nodes were packed, on the halo stream */
int chimp_set_exchange_callback(chimp_lattice *, chimp_exchange_fn fn, void *user);
/* N-rank twophase / std_one_phase need two more transports, supplied the same way:
 * - the scalar ghost exchange of cgField (MonLatMpi::communicateScalarField, LBmonlatmpi.h:181-205):
 *   one double per ghost node, packed in chimp_scalar_send_buffer_dev(k) -> peer's chimp_scalar_recv_buffer_dev;
 * - MPI_Allreduce(SUM) of `count` doubles in place at `dev` (momentum sum, main_TWOPHASE.cpp:299;
 *   mass change per interior domain, std_one_phase/main.cpp:528). */
typedef int (*chimp_allreduce_fn)(void *user, double *dev, int count, void *stream);
int chimp_set_scalar_exchange_callback(chimp_lattice *, chimp_exchange_fn fn, void *user);
int chimp_set_allreduce_callback(chimp_lattice *, chimp_allreduce_fn fn, void *user);
int chimp_scalar_neighbor_info(chimp_lattice *, int k, long long *send_count, long long *recv_count);
void *chimp_scalar_send_buffer_dev(chimp_lattice *, int k);
void *chimp_scalar_recv_buffer_dev(chimp_lattice *, int k);
/* use an externally owned stream (e.g. torch's current stream) for all launches */
int chimp_set_stream(chimp_lattice *, void *cuda_stream);
int chimp_synchronize(chimp_lattice *);

/* ---- introspection for tests / bench */
int chimp_num_own_nodes(chimp_lattice *);
/* fraction of (tile, q) pairs that needed explicit rows in CHIMP_INDEX_COMPACT form */
double chimp_irregular_fraction(chimp_lattice *);
/* device bytes of index data read per node per step, and of population data */
double chimp_index_bytes_per_node(chimp_lattice *);
/* Compact index, plain single-field step (no one_phase attributes, outside the fused peer launch): on != 0 selects the
 * kernel form that first reads a tile's bases and skips every delta word the tile's skip mask marks as a plain run
 * (byte == lane: the sources of those four directions are 32 consecutive slots).  Same tables, same results; fewer index
 * bytes (chimp_index_bytes_per_node follows) behind one more dependent load.  Off by default; bench.py tries both forms
 * before it times.  chimp_index_skipped_word_fraction: share of the (tile, delta word) pairs that are such runs. */
int chimp_set_index_skip_mask(chimp_lattice *, int on);
double chimp_index_skipped_word_fraction(chimp_lattice *);
/* one_phase lattices: bytes per node and step of the per-node attributes the step kernel reads (force switch, source
 * switch, interior label, link mask): 4 when they pack into one word (switches exactly 0 / 1, at most 16 labels;
 * CHIMP_ATTR_PACKED=0 keeps the arrays), 24 as four arrays, 0 before chimp_set_one_phase_attributes */
double chimp_one_phase_attribute_bytes_per_node(chimp_lattice *);
/* two-field lattices: device bytes per node per step of the index that locates phi of neighbor(q, n) for the colour
 * gradient (LButilities.h:12-22): 4 nQ for the table, 4 + 4 (stored words) / n for the derived form (the slot is
 * the pull source of rev(q) unless the link is stored; CHIMP_PHI_DERIVED=0 keeps the table); 0 without a phi table */
double chimp_phi_index_bytes_per_node(chimp_lattice *);
/* device pointer / geometry of the population planes (for bench timing of the bare kernel) */
long long chimp_plane_stride(chimp_lattice *);

#ifdef __cplusplus
}
#endif
#endif /* CHIMP_B200_H */
