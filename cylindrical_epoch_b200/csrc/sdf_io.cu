// sdf_io.cu -- SDF dump / restart of the hot-path state straight from the device mirrors
// (SURVEY.md section 8(f)4): the 30 mode-array blocks the reference writes with write_mode_field
// (io/diagnostics.F90:497-575,2033-2110), the field grid (:869-873) and, per species, the particle
// grid and the Weight / Px / Py / Pz point variables (:3040-3073,3110-3160; :614,:695-699), in the
// reference's own container: SDF 1.4 (SDF/FORTRAN/src/sdf_output*.f90 writes it, SDF/C reads it).
//
// File layout restated from the format the reference's library defines (no library code is used):
//   file header   "SDF1", endianness 16911887, version 1, revision 4, code name[32], first block
//                 location, summary location, summary size, nblocks, block header length, step,
//                 time, jobid1, jobid2, string length, code io version, restart flag, subdomain
//                 flag, station flag, 5 bytes of padding      (sdf_control.h:20, sdf_output.c:242-318)
//   block header  next block location, data location, id[32], data length, block type, data type,
//                 ndims, name[string length], length of the block's own metadata   (:323-370)
//   plain mesh    mults[nd], labels[nd][32], units[nd][32], geometry, min[nd], max[nd], dims[nd]   (:776-836)
//   plain var     mult, units[32], mesh id[32], dims[nd], stagger                                  (:841-890)
//   point mesh    mults[nd], labels, units, geometry, min[nd], max[nd], npoints (i8), species id   (:895-943)
//   point var     mult, units[32], mesh id[32], npoints (i8), species id[32]                       (:948-986)
// Metadata is written inline (no summary: summary size 0, as sdf_output.c:256-258 starts a file).
//
// Every rank of an x-slab run writes its own pieces into the same file at offsets that follow from
// the global sizes alone (plus the particle offsets the caller supplies, the reference's
// species_offset), so no rank waits for another: rank 0 adds the metadata and sets the file size.
// The host-level entry points (cylgpu_sdf_write_host / cylgpu_sdf_read_host) work on host arrays and
// need no GPU; cylgpu_sdf_dump / cylgpu_sdf_load move the data between the file and the device.
// Product code: never includes, links or calls anything under oracle/.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "ctx.cuh"

namespace cylgpu {

namespace {

constexpr int ID_LEN = 32, STR_LEN = 64;
constexpr int64_t FILE_HEADER_LEN = 11 * 4 + 2 * 8 + 8 + 12 + ID_LEN;           // sdf_control.h:20
constexpr int32_t BLOCK_HEADER_LEN = 4 + 3 * 4 + 3 * 8 + ID_LEN + STR_LEN;      // sdf_control.h:21-22
constexpr int32_t SDF_ENDIAN = 16911887, SDF_VER = 1, SDF_REV = 4;
enum { BT_PLAIN_MESH = 1, BT_POINT_MESH = 2, BT_PLAIN_VARIABLE = 3, BT_POINT_VARIABLE = 4, BT_CONSTANT = 5 };   // sdf.h:47-61
constexpr int32_t DT_REAL8 = 4;                                                   // sdf.h:195-200
constexpr int32_t GEOMETRY_CARTESIAN = 1;                                         // sdf.h:130-132
// sdf_common.f90:184-195; constants.F90:291-299
constexpr int32_t STAG_FACE_X = 1, STAG_FACE_Y = 2, STAG_EDGE_Z = 3, STAG_FACE_Z = 4;

struct ModeBlock { const char* stem; const char* group; const char* comp; const char* units; int field; int32_t stagger; };
// the order of io/diagnostics.F90:497-575: per group the three real parts, then the three imaginary parts
const ModeBlock MODE_GROUPS[5][3] = {
    {{"exm", "Electric Field Modes", "Exm", "V/m", CYLGPU_EXM, STAG_FACE_Y},
     {"erm", "Electric Field Modes", "Erm", "V/m", CYLGPU_ERM, STAG_FACE_X},
     {"etm", "Electric Field Modes", "Etm", "V/m", CYLGPU_ETM, STAG_EDGE_Z}},
    {{"bxm", "Magnetic Field Modes", "Bxm", "T", CYLGPU_BXM, STAG_FACE_X},
     {"brm", "Magnetic Field Modes", "Brm", "T", CYLGPU_BRM, STAG_FACE_Y},
     {"btm", "Magnetic Field Modes", "Btm", "T", CYLGPU_BTM, STAG_FACE_Z}},
    {{"bxm_old", "Magnetic Field Modes", "Bxm_old", "T", CYLGPU_BXM_OLD, STAG_FACE_X},
     {"brm_old", "Magnetic Field Modes", "Brm_old", "T", CYLGPU_BRM_OLD, STAG_FACE_Y},
     {"btm_old", "Magnetic Field Modes", "Btm_old", "T", CYLGPU_BTM_OLD, STAG_FACE_Z}},
    {{"jxm", "Current Modes", "Jxm", "A/m^2", CYLGPU_JXM, STAG_FACE_Y},
     {"jrm", "Current Modes", "Jrm", "A/m^2", CYLGPU_JRM, STAG_FACE_X},
     {"jtm", "Current Modes", "Jtm", "A/m^2", CYLGPU_JTM, STAG_EDGE_Z}},
    {{"jxm_old", "Current Modes", "Jxm_old", "A/m^2", CYLGPU_JXM_OLD, STAG_FACE_Y},
     {"jrm_old", "Current Modes", "Jrm_old", "A/m^2", CYLGPU_JRM_OLD, STAG_FACE_X},
     {"jtm_old", "Current Modes", "Jtm_old", "A/m^2", CYLGPU_JTM_OLD, STAG_EDGE_Z}},
};
// r-staggered arrays are written shifted by one row so that file row 1 is the axis row 0
// (io/diagnostics.F90:2085-2097: stagger == c_stagger_exm .OR. c_stagger_etm)
inline int row_shift(int32_t stagger) { return (stagger == STAG_FACE_Y || stagger == STAG_EDGE_Z) ? 1 : 0; }

struct PointVar { const char* name; const char* units; int comp; };   // comp: index into the 7-double record
const PointVar POINT_VARS[4] = {{"Weight", "", 6}, {"Px", "kg.m/s", 3}, {"Py", "kg.m/s", 4}, {"Pz", "kg.m/s", 5}};

// write_nspecies_field call sites, io/diagnostics.F90:765-835 (poynt_flux reads the legacy Cartesian arrays: not offered)
struct DerivedVar { const char* id; const char* name; const char* units; };
const DerivedVar DERIVED[CYLGPU_SDF_NDERIVED] = {
    {"ekbar", "Average_Particle_Energy", "J"}, {"mass_density", "Mass_Density", "kg/m^3"},
    {"charge_density", "Charge_Density", "C/m^3"}, {"number_density", "Number_Density", "1/m^3"},
    {"ppc", "Particles_Per_Cell", "n_particles"}, {"average_weight", "Particles_Average_Weight", "weight"},
    {"average_px", "Particles_Average_Px", "kg.m/s"}, {"average_py", "Particles_Average_Py", "kg.m/s"},
    {"average_pz", "Particles_Average_Pz", "kg.m/s"}, {"temperature", "Temperature", "K"},
    {"temperature_x", "Temperature_x", "K"}, {"temperature_y", "Temperature_y", "K"},
    {"temperature_z", "Temperature_z", "K"}, {"jx", "Jx", "A/m^2"}, {"jy", "Jy", "A/m^2"}, {"jz", "Jz", "A/m^2"},
    {"ekflux/x_max", "Particle_Energy_Flux/x_max", "W/m^2"}, {"ekflux/y_max", "Particle_Energy_Flux/y_max", "W/m^2"},
    {"ekflux/z_max", "Particle_Energy_Flux/z_max", "W/m^2"}, {"ekflux/x_min", "Particle_Energy_Flux/x_min", "W/m^2"},
    {"ekflux/y_min", "Particle_Energy_Flux/y_min", "W/m^2"}, {"ekflux/z_min", "Particle_Energy_Flux/z_min", "W/m^2"}};

struct Bytes {
  std::vector<unsigned char> b;
  template <class T> void put(T v) { const unsigned char* p = reinterpret_cast<const unsigned char*>(&v); b.insert(b.end(), p, p + sizeof(T)); }
  // blank-trimmed, NUL-padded fixed-length field.  Like the Fortran writer (sdf_safe_copy_id / _string,
  // SDF/FORTRAN/src/sdf_common.f90) a name may fill all `len` characters -- longer ones are cut there,
  // e.g. 'number_density_mode/electron/Real' is stored as its first 32 characters.
  void str(const std::string& s, int len) {
    size_t a = 0, e = s.size();
    while (a < e && isspace((unsigned char)s[a])) ++a;
    while (e > a && isspace((unsigned char)s[e - 1])) --e;
    for (int i = 0; i < len; ++i) b.push_back((size_t)i < e - a ? (unsigned char)s[a + i] : 0);
  }
};

struct Block {
  std::string id, name;
  int32_t blocktype, ndims;
  Bytes meta;              // the block-type specific metadata
  int64_t data_length;
  int64_t start = 0, data_location = 0, next = 0;
};

std::string lower(std::string s) { for (char& c : s) c = (char)tolower((unsigned char)c); return s; }

int pwrite_all(int fd, const void* buf, size_t n, int64_t off) {
  const char* p = static_cast<const char*>(buf);
  while (n > 0) {
    const ssize_t w = pwrite(fd, p, n, (off_t)off);
    if (w < 0) { if (errno == EINTR) continue; set_error("sdf: write failed: %s", strerror(errno)); return 1; }
    p += w; off += w; n -= (size_t)w;
  }
  return 0;
}
int pread_all(int fd, void* buf, size_t n, int64_t off) {
  char* p = static_cast<char*>(buf);
  while (n > 0) {
    const ssize_t r = pread(fd, p, n, (off_t)off);
    if (r < 0) { if (errno == EINTR) continue; set_error("sdf: read failed: %s", strerror(errno)); return 1; }
    if (r == 0) { set_error("sdf: unexpected end of file"); return 1; }
    p += r; off += r; n -= (size_t)r;
  }
  return 0;
}

int check_desc(const cylgpu_sdf_desc* d) {
  if (!d || d->nx_global < 1 || d->ny_global < 1 || d->n_mode < 1 || d->nx_local < 1 || d->cell_x_min < 1 ||
      d->cell_x_min + d->nx_local - 1 > d->nx_global || d->n_species < 0 || d->n_species > CYLGPU_MAX_SPECIES) {
    set_error("sdf: bad descriptor");
    return 2;
  }
  if (d->n_constants < 0 || d->n_constants > CYLGPU_SDF_MAX_CONSTANTS) { set_error("sdf: bad number of constants"); return 2; }
  for (int k = 0; k < d->n_constants; ++k)
    if (!d->constant_id[k] || !d->constant_id[k][0]) { set_error("sdf: constant %d has no id", k); return 2; }
  for (int s = 0; s < d->n_species; ++s) {
    if (!d->species_name[s] || !d->species_name[s][0]) { set_error("sdf: species %d has no name", s); return 2; }
    if (d->npart_local[s] < 0 || d->npart_offset[s] < 0 || d->npart_offset[s] + d->npart_local[s] > d->npart_global[s]) {
      set_error("sdf: species %d: local particles [%lld, +%lld) do not fit the global count %lld", s,
                (long long)d->npart_offset[s], (long long)d->npart_local[s], (long long)d->npart_global[s]);
      return 2;
    }
  }
  return 0;
}

// the block list of a dump: offsets follow from the descriptor alone, identically on every rank
std::vector<Block> build_blocks(const cylgpu_sdf_desc* d) {
  std::vector<Block> bl;
  const int nxg = d->nx_global, nyg = d->ny_global, M = d->n_mode;
  for (int k = 0; k < d->n_constants; ++k) {   // sdf_write_srl(id, name, value): the value is the block's metadata
    Block b;
    b.id = d->constant_id[k];
    b.name = d->constant_name[k] ? d->constant_name[k] : d->constant_id[k];
    b.blocktype = BT_CONSTANT; b.ndims = 1;
    b.meta.put<double>(d->constant_value[k]);
    b.data_length = 0;
    bl.push_back(b);
  }
  {   // sdf_write_srl_plain_mesh('grid', 'Grid/Grid', xb_global, yb_global), io/diagnostics.F90:872
    Block b;
    b.id = "grid"; b.name = "Grid/Grid"; b.blocktype = BT_PLAIN_MESH; b.ndims = 2;
    b.meta.put<double>(1.0); b.meta.put<double>(1.0);
    b.meta.str("X", ID_LEN); b.meta.str("Y", ID_LEN);
    b.meta.str("m", ID_LEN); b.meta.str("m", ID_LEN);
    b.meta.put<int32_t>(GEOMETRY_CARTESIAN);
    b.meta.put<double>(d->x_min); b.meta.put<double>(0.0);
    b.meta.put<double>(d->x_min + nxg * d->dx); b.meta.put<double>(nyg * d->dy);
    b.meta.put<int32_t>(nxg + 1); b.meta.put<int32_t>(nyg + 1);
    b.data_length = (int64_t)(nxg + 1 + nyg + 1) * 8;
    bl.push_back(b);
  }
  for (int grp = 0; grp < 5; ++grp)
    for (int part = 0; part < 2; ++part)
      for (int k = 0; k < 3; ++k) {
        const ModeBlock& mb = MODE_GROUPS[grp][k];
        Block b;
        b.id = std::string(mb.stem) + (part ? "_imag" : "_real");
        b.name = std::string(mb.group) + "/" + mb.comp + (part ? "/imag" : "/real");
        b.blocktype = BT_PLAIN_VARIABLE; b.ndims = 3;
        b.meta.put<double>(1.0);
        b.meta.str(mb.units, ID_LEN);
        b.meta.str("mode_grid", ID_LEN);      // io/diagnostics.F90:2105
        b.meta.put<int32_t>(nxg); b.meta.put<int32_t>(nyg); b.meta.put<int32_t>(M);
        b.meta.put<int32_t>(mb.stagger);
        b.data_length = (int64_t)nxg * nyg * M * 8;
        bl.push_back(b);
      }
  for (int s = 0; s < d->n_species; ++s) {   // write_particle_grid, io/diagnostics.F90:3051-3073
    if (d->npart_global[s] == 0) continue;
    Block b;
    b.id = std::string("grid/") + d->species_name[s];
    b.name = std::string("Grid/Particles/") + d->species_name[s];
    b.blocktype = BT_POINT_MESH; b.ndims = 3;
    for (int k = 0; k < 3; ++k) b.meta.put<double>(1.0);
    b.meta.str("X", ID_LEN); b.meta.str("Y", ID_LEN); b.meta.str("Z", ID_LEN);
    for (int k = 0; k < 3; ++k) b.meta.str("m", ID_LEN);
    b.meta.put<int32_t>(GEOMETRY_CARTESIAN);
    for (int k = 0; k < 6; ++k) b.meta.put<double>(d->part_extents[s][k]);
    b.meta.put<int64_t>(d->npart_global[s]);
    b.meta.str(d->species_name[s], ID_LEN);
    b.data_length = 3 * d->npart_global[s] * 8;
    bl.push_back(b);
  }
  for (int v = 0; v < 4; ++v)
    for (int s = 0; s < d->n_species; ++s) {   // write_particle_variable, io/diagnostics.F90:3133-3160
      if (d->npart_global[s] == 0) continue;
      Block b;
      b.id = lower(std::string(POINT_VARS[v].name) + "/" + d->species_name[s]);
      b.name = std::string("Particles/") + POINT_VARS[v].name + "/" + d->species_name[s];
      b.blocktype = BT_POINT_VARIABLE; b.ndims = 1;
      b.meta.put<double>(1.0);
      b.meta.str(POINT_VARS[v].units, ID_LEN);
      b.meta.str(std::string("grid/") + d->species_name[s], ID_LEN);
      b.meta.put<int64_t>(d->npart_global[s]);
      b.meta.str(d->species_name[s], ID_LEN);
      b.data_length = d->npart_global[s] * 8;
      bl.push_back(b);
    }
  for (int v = 0; v < CYLGPU_SDF_NDERIVED; ++v) {   // write_nspecies_field, io/diagnostics.F90:2222-2232,2396-2431
    if (!(d->derived_mask & (1u << v))) continue;
    for (int s = d->derived_sum ? -1 : 0; s < (d->derived_species ? d->n_species : 0); ++s) {
      Block b;
      b.id = DERIVED[v].id;
      b.name = std::string("Derived/") + DERIVED[v].name;
      if (s >= 0) { b.id += std::string("/") + d->species_name[s]; b.name += std::string("/") + d->species_name[s]; }
      b.blocktype = BT_PLAIN_VARIABLE; b.ndims = 2;
      b.meta.put<double>(1.0);
      b.meta.str(DERIVED[v].units, ID_LEN);
      b.meta.str("grid", ID_LEN);
      b.meta.put<int32_t>(nxg); b.meta.put<int32_t>(nyg);
      b.meta.put<int32_t>(0);   // c_stagger_cell_centre
      b.data_length = (int64_t)nxg * nyg * 8;
      bl.push_back(b);
    }
  }
  if (d->derived_mask & (1u << CYLGPU_SDF_NDERIVED))   // write_nspecies_field_mode, io/diagnostics.F90:2596-2676
    for (int s = 0; s < d->n_species; ++s)
      for (int part = 0; part < 2; ++part) {
        Block b;
        const std::string tail = std::string("/") + d->species_name[s] + (part ? "/Imaginary" : "/Real");
        b.id = "number_density_mode" + tail;
        b.name = "Number_Density_Mode" + tail;
        b.blocktype = BT_PLAIN_VARIABLE; b.ndims = 3;
        b.meta.put<double>(1.0);
        b.meta.str("1/m^3", ID_LEN);
        b.meta.str("mode_grid", ID_LEN);
        b.meta.put<int32_t>(nxg); b.meta.put<int32_t>(nyg); b.meta.put<int32_t>(M);
        b.meta.put<int32_t>(0);   // c_stagger_cell_centre
        b.data_length = (int64_t)nxg * nyg * M * 8;
        bl.push_back(b);
      }
  int64_t pos = FILE_HEADER_LEN;
  for (Block& b : bl) {
    if (b.id.size() > (size_t)ID_LEN) b.id.resize(ID_LEN);
    b.start = pos;
    b.data_location = pos + BLOCK_HEADER_LEN + (int64_t)b.meta.b.size();
    b.next = b.data_location + b.data_length;
    pos = b.next;
  }
  return bl;
}

const Block* find_block(const std::vector<Block>& bl, std::string id) {
  if (id.size() > (size_t)ID_LEN) id.resize(ID_LEN);
  for (const Block& b : bl) if (b.id == id) return &b;
  return nullptr;
}

struct Fd {
  int fd = -1;
  ~Fd() { if (fd >= 0) close(fd); }
};

}  // namespace

// fields15[id]: host array of field id (include/cylgpu.h CYLGPU_EXM ..), complex(num)
// (1-ng:nx_local+ng, 1-ng:ny+ng, 0:M-1); particles_aos[s]: npart_local[s] records of 7 doubles.
int sdf_derived_count(const cylgpu_sdf_desc* d) {
  int n = 0;
  for (int v = 0; v < CYLGPU_SDF_NDERIVED; ++v)
    if (d->derived_mask & (1u << v)) n += (d->derived_sum ? 1 : 0) + (d->derived_species ? d->n_species : 0);
  if (d->derived_mask & (1u << CYLGPU_SDF_NDERIVED)) n += d->n_species;   // number_density_mode: one complex array each
  return n;
}

int sdf_write_host(const char* path, const cylgpu_sdf_desc* d, const void* const* fields15,
                   const double* const* particles_aos, const double* const* derived) {
  TRY(check_desc(d));
  if (!path || !fields15) { set_error("sdf: null argument"); return 2; }
  const std::vector<Block> bl = build_blocks(d);
  const bool master = d->cell_x_min == 1;   // the rank that owns x_min writes the metadata
  Fd f;
  f.fd = open(path, O_WRONLY | O_CREAT, 0644);
  if (f.fd < 0) { set_error("sdf: cannot open %s: %s", path, strerror(errno)); return 1; }
  const int nxg = d->nx_global, nyg = d->ny_global, M = d->n_mode, nxl = d->nx_local;
  if (master) {
    if (ftruncate(f.fd, (off_t)bl.back().next) != 0) { set_error("sdf: ftruncate: %s", strerror(errno)); return 1; }
    Bytes h;
    h.b.insert(h.b.end(), {'S', 'D', 'F', '1'});
    h.put<int32_t>(SDF_ENDIAN); h.put<int32_t>(SDF_VER); h.put<int32_t>(SDF_REV);
    h.str("Epoch2d", ID_LEN);                           // io/diagnostics.F90:393
    h.put<int64_t>(FILE_HEADER_LEN);                    // first block location
    h.put<int64_t>(bl.back().next);                     // summary location: end of file, none written
    h.put<int32_t>(0);                                  // summary size
    h.put<int32_t>((int32_t)bl.size());
    h.put<int32_t>(BLOCK_HEADER_LEN);
    h.put<int32_t>(d->step);
    h.put<double>(d->time);
    h.put<int32_t>(d->jobid1); h.put<int32_t>(d->jobid2);
    h.put<int32_t>(STR_LEN);
    h.put<int32_t>(1);                                  // c_code_io_version, version_data.F90:23
    h.b.push_back(d->restart ? 1 : 0);
    h.b.push_back(0);                                   // subdomain file
    h.b.push_back(0);                                   // station file
    for (int i = 0; i < 5; ++i) h.b.push_back(0);
    if ((int64_t)h.b.size() != FILE_HEADER_LEN) { set_error("sdf: internal header length"); return 1; }
    TRY(pwrite_all(f.fd, h.b.data(), h.b.size(), 0));
    for (const Block& b : bl) {
      Bytes bh;
      bh.put<int64_t>(b.next); bh.put<int64_t>(b.data_location);
      bh.str(b.id, ID_LEN);
      bh.put<int64_t>(b.data_length);
      bh.put<int32_t>(b.blocktype); bh.put<int32_t>(DT_REAL8); bh.put<int32_t>(b.ndims);
      bh.str(b.name, STR_LEN);
      bh.put<int32_t>((int32_t)b.meta.b.size());
      if ((int32_t)bh.b.size() != BLOCK_HEADER_LEN) { set_error("sdf: internal block header length"); return 1; }
      bh.b.insert(bh.b.end(), b.meta.b.begin(), b.meta.b.end());
      TRY(pwrite_all(f.fd, bh.b.data(), bh.b.size(), b.start));
    }
    std::vector<double> xy((size_t)nxg + 1 + nyg + 1);
    for (int i = 0; i <= nxg; ++i) xy[i] = d->x_min + (double)i * d->dx;          // xb_global(1:nx_global+1)
    for (int j = 0; j <= nyg; ++j) xy[(size_t)nxg + 1 + j] = (double)j * d->dy;   // yb_global(1:ny_global+1)
    TRY(pwrite_all(f.fd, xy.data(), xy.size() * 8, find_block(bl, "grid")->data_location));
  }
  // mode arrays: file element (ix_global, j, im), j = 1..ny, holds array row j - shift
  const int SX = nxl + 2 * NG, SY = nyg + 2 * NG;
  std::vector<double> slab((size_t)nxl * nyg * M);
  for (int grp = 0; grp < 5; ++grp)
    for (int part = 0; part < 2; ++part)
      for (int k = 0; k < 3; ++k) {
        const ModeBlock& mb = MODE_GROUPS[grp][k];
        const double* a = static_cast<const double*>(fields15[mb.field]);
        if (!a) { set_error("sdf: field %s missing", mb.stem); return 2; }
        const int sh = row_shift(mb.stagger);
        for (int im = 0; im < M; ++im)
          for (int j = 1; j <= nyg; ++j) {
            const double* src = a + 2 * (((size_t)im * SY + (size_t)(j - sh + NG - 1)) * SX + NG) + part;
            double* dst = slab.data() + ((size_t)im * nyg + (j - 1)) * nxl;
            for (int i = 0; i < nxl; ++i) dst[i] = src[2 * (size_t)i];
          }
        const Block* b = find_block(bl, std::string(mb.stem) + (part ? "_imag" : "_real"));
        if (nxl == nxg) {
          TRY(pwrite_all(f.fd, slab.data(), slab.size() * 8, b->data_location));
        } else {
          for (int im = 0; im < M; ++im)
            for (int j = 0; j < nyg; ++j)
              TRY(pwrite_all(f.fd, slab.data() + ((size_t)im * nyg + j) * nxl, (size_t)nxl * 8,
                             b->data_location + (((int64_t)im * nyg + j) * nxg + (d->cell_x_min - 1)) * 8));
        }
      }
  // particles: component-major, this rank's records behind the npart_offset of the ranks before it
  for (int s = 0; s < d->n_species; ++s) {
    const int64_t nl = d->npart_local[s], ng_ = d->npart_global[s];
    if (ng_ == 0 || nl == 0) continue;
    if (!particles_aos || !particles_aos[s]) { set_error("sdf: particle list of species %d missing", s); return 2; }
    const double* p = particles_aos[s];
    std::vector<double> col((size_t)nl);
    const Block* gm = find_block(bl, std::string("grid/") + d->species_name[s]);
    for (int c = 0; c < 3; ++c) {
      for (int64_t i = 0; i < nl; ++i) col[(size_t)i] = p[7 * i + c];
      TRY(pwrite_all(f.fd, col.data(), (size_t)nl * 8, gm->data_location + ((int64_t)c * ng_ + d->npart_offset[s]) * 8));
    }
    for (int v = 0; v < 4; ++v) {
      const Block* b = find_block(bl, lower(std::string(POINT_VARS[v].name) + "/" + d->species_name[s]));
      for (int64_t i = 0; i < nl; ++i) col[(size_t)i] = p[7 * i + POINT_VARS[v].comp];
      TRY(pwrite_all(f.fd, col.data(), (size_t)nl * 8, b->data_location + d->npart_offset[s] * 8));
    }
  }
  // derived variables: the interior of a cell-centred real array per block
  if (sdf_derived_count(d) > 0) {
    if (!derived) { set_error("sdf: derived variables selected but no arrays given"); return 2; }
    int q = 0;
    std::vector<double> rows((size_t)nxl * nyg);
    for (int v = 0; v < CYLGPU_SDF_NDERIVED; ++v) {
      if (!(d->derived_mask & (1u << v))) continue;
      for (int s = d->derived_sum ? -1 : 0; s < (d->derived_species ? d->n_species : 0); ++s, ++q) {
        const double* a = derived[q];
        if (!a) { set_error("sdf: derived array %d missing", q); return 2; }
        std::string id = DERIVED[v].id;
        if (s >= 0) id += std::string("/") + d->species_name[s];
        const Block* b = find_block(bl, id);
        for (int j = 1; j <= nyg; ++j)
          memcpy(rows.data() + (size_t)(j - 1) * nxl, a + (size_t)(j + NG - 1) * SX + NG, (size_t)nxl * 8);
        if (nxl == nxg) {
          TRY(pwrite_all(f.fd, rows.data(), rows.size() * 8, b->data_location));
        } else {
          for (int j = 0; j < nyg; ++j)
            TRY(pwrite_all(f.fd, rows.data() + (size_t)j * nxl, (size_t)nxl * 8,
                           b->data_location + ((int64_t)j * nxg + (d->cell_x_min - 1)) * 8));
        }
      }
    }
  }
  if (d->derived_mask & (1u << CYLGPU_SDF_NDERIVED)) {
    // the per-species density modes: complex arrays (1-ng:nx+ng, 1-ng:ny+ng, 0:M-1) behind the real-valued ones
    int q = sdf_derived_count(d) - d->n_species;
    for (int s = 0; s < d->n_species; ++s, ++q) {
      const double* a = derived ? derived[q] : nullptr;
      if (!a) { set_error("sdf: number_density_mode array of species %d missing", s); return 2; }
      for (int part = 0; part < 2; ++part) {
        for (int im = 0; im < M; ++im)
          for (int j = 1; j <= nyg; ++j) {
            const double* src = a + 2 * (((size_t)im * SY + (size_t)(j + NG - 1)) * SX + NG) + part;
            double* dst = slab.data() + ((size_t)im * nyg + (j - 1)) * nxl;
            for (int i = 0; i < nxl; ++i) dst[i] = src[2 * (size_t)i];
          }
        const Block* b = find_block(bl, std::string("number_density_mode/") + d->species_name[s] + (part ? "/Imaginary" : "/Real"));
        if (nxl == nxg) {
          TRY(pwrite_all(f.fd, slab.data(), slab.size() * 8, b->data_location));
        } else {
          for (int im = 0; im < M; ++im)
            for (int j = 0; j < nyg; ++j)
              TRY(pwrite_all(f.fd, slab.data() + ((size_t)im * nyg + j) * nxl, (size_t)nxl * 8,
                             b->data_location + (((int64_t)im * nyg + j) * nxg + (d->cell_x_min - 1)) * 8));
        }
      }
    }
  }
  if (fsync(f.fd) != 0) { set_error("sdf: fsync: %s", strerror(errno)); return 1; }
  return 0;
}

// ---- reader of the same subset (restart of the hot-path state, housekeeping/setup.F90:1196-1260,1424-1466) ----
namespace {

struct FileBlock { std::string id; int64_t data_location, data_length; int32_t blocktype, datatype, ndims; std::vector<unsigned char> meta; };

int read_blocklist(int fd, std::vector<FileBlock>& out, int32_t* step, double* time) {
  unsigned char h[FILE_HEADER_LEN];
  TRY(pread_all(fd, h, sizeof h, 0));
  int32_t endian, ver, nblocks, bhl, strl;
  int64_t first;
  memcpy(&endian, h + 4, 4); memcpy(&ver, h + 8, 4);
  if (memcmp(h, "SDF1", 4) != 0 || endian != SDF_ENDIAN || ver != SDF_VER) { set_error("sdf: not an SDF 1.x file of this endianness"); return 2; }
  memcpy(&first, h + 16 + ID_LEN, 8);
  memcpy(&nblocks, h + 16 + ID_LEN + 20, 4);
  memcpy(&bhl, h + 16 + ID_LEN + 24, 4);
  memcpy(step, h + 16 + ID_LEN + 28, 4);
  memcpy(time, h + 16 + ID_LEN + 32, 8);
  memcpy(&strl, h + 16 + ID_LEN + 48, 4);
  if (bhl != 4 + 3 * 4 + 3 * 8 + ID_LEN + strl) { set_error("sdf: unexpected block header length %d", bhl); return 2; }
  int64_t pos = first;
  for (int n = 0; n < nblocks; ++n) {
    std::vector<unsigned char> bh((size_t)bhl);
    TRY(pread_all(fd, bh.data(), bh.size(), pos));
    FileBlock b;
    int64_t next;
    int32_t info;
    memcpy(&next, bh.data(), 8);
    memcpy(&b.data_location, bh.data() + 8, 8);
    b.id.assign(reinterpret_cast<const char*>(bh.data() + 16), strnlen(reinterpret_cast<const char*>(bh.data() + 16), ID_LEN));
    memcpy(&b.data_length, bh.data() + 16 + ID_LEN, 8);
    memcpy(&b.blocktype, bh.data() + 24 + ID_LEN, 4);
    memcpy(&b.datatype, bh.data() + 28 + ID_LEN, 4);
    memcpy(&b.ndims, bh.data() + 32 + ID_LEN, 4);
    memcpy(&info, bh.data() + 36 + ID_LEN + strl, 4);
    if (info < 0 || info > (1 << 20)) { set_error("sdf: corrupt block %d", n); return 2; }
    b.meta.resize((size_t)info);
    if (info > 0) TRY(pread_all(fd, b.meta.data(), b.meta.size(), pos + bhl));
    out.push_back(b);
    if (next <= pos) break;
    pos = next;
  }
  return 0;
}

const FileBlock* find_file_block(const std::vector<FileBlock>& bl, const std::string& id) {
  for (const FileBlock& b : bl) if (b.id == id) return &b;
  return nullptr;
}

}  // namespace

// Reads this rank's slab of the 15 mode arrays into host arrays with ghosts (interior rows and columns only,
// the shift of the r-staggered arrays undone: setup.F90:1199-1210) and the particles with
// x_lo <= x < x_hi of every named species.  The ghosts stay as they are in the destination (the
// caller zeroes them and re-derives them with the boundary routines, as the reference's restart does).
int sdf_read_host(const char* path, cylgpu_sdf_desc* d, void* const* fields15, double x_lo, double x_hi,
                  std::vector<std::vector<double>>* particles) {
  if (!path || !d || !fields15) { set_error("sdf: null argument"); return 2; }
  Fd f;
  f.fd = open(path, O_RDONLY);
  if (f.fd < 0) { set_error("sdf: cannot open %s: %s", path, strerror(errno)); return 1; }
  std::vector<FileBlock> bl;
  TRY(read_blocklist(f.fd, bl, &d->step, &d->time));
  d->constants_found = 0;
  for (int k = 0; k < d->n_constants && k < CYLGPU_SDF_MAX_CONSTANTS; ++k) {
    const FileBlock* b = d->constant_id[k] ? find_file_block(bl, d->constant_id[k]) : nullptr;
    if (b && b->blocktype == BT_CONSTANT && b->datatype == DT_REAL8 && b->meta.size() >= 8) {
      memcpy(&d->constant_value[k], b->meta.data(), 8);
      d->constants_found |= 1u << k;
    }
  }
  const int nxg = d->nx_global, nyg = d->ny_global, M = d->n_mode, nxl = d->nx_local;
  const int SX = nxl + 2 * NG, SY = nyg + 2 * NG;
  std::vector<double> row((size_t)nxl);
  for (int grp = 0; grp < 5; ++grp)
    for (int part = 0; part < 2; ++part)
      for (int k = 0; k < 3; ++k) {
        const ModeBlock& mb = MODE_GROUPS[grp][k];
        const std::string id = std::string(mb.stem) + (part ? "_imag" : "_real");
        const FileBlock* b = find_file_block(bl, id);
        if (!b) { set_error("sdf: block %s not in %s", id.c_str(), path); return 2; }
        int32_t dims[3];
        if (b->blocktype != BT_PLAIN_VARIABLE || b->datatype != DT_REAL8 || b->ndims != 3 ||
            b->meta.size() < 8 + 2 * ID_LEN + 16) { set_error("sdf: block %s has an unexpected type", id.c_str()); return 2; }
        memcpy(dims, b->meta.data() + 8 + 2 * ID_LEN, 12);
        if (dims[0] != nxg || dims[1] != nyg || dims[2] != M) {
          set_error("sdf: block %s is %d x %d x %d, expected %d x %d x %d", id.c_str(), dims[0], dims[1], dims[2], nxg, nyg, M);
          return 2;
        }
        double* a = static_cast<double*>(fields15[mb.field]);
        const int sh = row_shift(mb.stagger);
        for (int im = 0; im < M; ++im)
          for (int j = 1; j <= nyg; ++j) {
            TRY(pread_all(f.fd, row.data(), (size_t)nxl * 8,
                          b->data_location + (((int64_t)im * nyg + (j - 1)) * nxg + (d->cell_x_min - 1)) * 8));
            double* dst = a + 2 * (((size_t)im * SY + (size_t)(j - sh + NG - 1)) * SX + NG) + part;
            for (int i = 0; i < nxl; ++i) dst[2 * (size_t)i] = row[(size_t)i];
          }
      }
  if (!particles) return 0;
  particles->assign((size_t)d->n_species, std::vector<double>());
  for (int s = 0; s < d->n_species; ++s) {
    d->npart_global[s] = d->npart_local[s] = d->npart_offset[s] = 0;
    const FileBlock* gm = find_file_block(bl, std::string("grid/") + d->species_name[s]);
    if (!gm) continue;   // a species without particles has no blocks (io/diagnostics.F90:3062)
    if (gm->blocktype != BT_POINT_MESH || gm->ndims != 3 || gm->meta.size() < (size_t)(3 * 8 + 6 * ID_LEN + 4 + 48 + 8)) {
      set_error("sdf: particle grid of %s has an unexpected type", d->species_name[s]);
      return 2;
    }
    int64_t npg;
    memcpy(&npg, gm->meta.data() + 3 * 8 + 6 * ID_LEN + 4 + 48, 8);
    d->npart_global[s] = npg;
    const FileBlock* vb[4];
    for (int v = 0; v < 4; ++v) {
      vb[v] = find_file_block(bl, lower(std::string(POINT_VARS[v].name) + "/" + d->species_name[s]));
      if (!vb[v]) { set_error("sdf: %s of %s not in %s", POINT_VARS[v].name, d->species_name[s], path); return 2; }
    }
    std::vector<double>& out = (*particles)[(size_t)s];
    const int64_t CH = 1 << 20;
    std::vector<double> buf((size_t)CH);
    std::vector<int64_t> mine;
    for (int64_t off = 0; off < npg; off += CH) {
      const int64_t m = std::min<int64_t>(CH, npg - off);
      TRY(pread_all(f.fd, buf.data(), (size_t)m * 8, gm->data_location + off * 8));
      mine.clear();
      for (int64_t i = 0; i < m; ++i) if (buf[(size_t)i] >= x_lo && buf[(size_t)i] < x_hi) mine.push_back(i);
      const size_t base = out.size();
      out.resize(base + 7 * mine.size());
      for (size_t q = 0; q < mine.size(); ++q) out[base + 7 * q] = buf[(size_t)mine[q]];
      for (int c = 1; c < 3; ++c) {
        TRY(pread_all(f.fd, buf.data(), (size_t)m * 8, gm->data_location + ((int64_t)c * npg + off) * 8));
        for (size_t q = 0; q < mine.size(); ++q) out[base + 7 * q + c] = buf[(size_t)mine[q]];
      }
      for (int v = 0; v < 4; ++v) {
        TRY(pread_all(f.fd, buf.data(), (size_t)m * 8, vb[v]->data_location + off * 8));
        for (size_t q = 0; q < mine.size(); ++q) out[base + 7 * q + POINT_VARS[v].comp] = buf[(size_t)mine[q]];
      }
    }
    d->npart_local[s] = (int64_t)(out.size() / 7);
  }
  return 0;
}

}  // namespace cylgpu
