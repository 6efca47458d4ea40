// Host-side flattened sparse tree: what src/vdb (VDB::openFile / loadBasic / loadExt, vdb/vdb.cpp:103-388) turns
// into a per-voxel sphere list in the reference becomes pointer-free root / internal / leaf tables plus a dense
// brick atlas that is staged once into HBM (BASELINE.json north_star).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vrs {

struct HostGrid {
  std::string name, grid_type;
  bool half = false;            // values were fp16 on disk
  bool level_set = false;       // grid class "level set" (cube.vdb) -> fog conversion at upload
  float background = 0.0f;
  double voxel_size = 1.0;
  double translation[3] = {0, 0, 0};
  uint32_t file_version = 0, compression = 0;

  std::vector<int32_t> root;          // 4 per entry: origin xyz, child (>=0 internal5 index, <0 ~tile)
  std::vector<int32_t> i5;            // [n5][32768]: >=0 internal4 index, <0 ~tile
  std::vector<int32_t> i4;            // [n4][4096]:  >=0 leaf index, <0 ~tile
  std::vector<float> tile_value;      // raw tile values; entry 0 = background (inactive)
  std::vector<uint8_t> tile_active;
  std::vector<int32_t> leaf_origin;   // 3 per leaf
  std::vector<uint64_t> leaf_mask;    // 8 words per leaf, bit n = offset (x<<6)|(y<<3)|z
  std::vector<float> leaf_value;      // 512 per leaf, raw (inactive voxels filled by the file's rule)

  int32_t bbox_min[3] = {0, 0, 0}, bbox_max[3] = {-1, -1, -1};
  uint64_t active_voxels = 0;
  uint32_t root_children = 0;

  size_t n5() const { return i5.size() / 32768; }
  size_t n4() const { return i4.size() / 4096; }
  size_t nleaf() const { return leaf_origin.size() / 3; }

  int32_t add_tile(float value, bool active);
  void finalize();                                   // bbox + active voxel count
  // ValueAccessor::getValue semantics (vdb/vdb.cpp:777-786): raw value, active flag
  float get_value(int32_t i, int32_t j, int32_t k, bool* active) const;
  float density_from_raw(float raw) const;
};

bool read_vdb(const std::string& path, const char* grid_name, HostGrid& out, std::string& err);
bool write_vrsg(const std::string& path, const HostGrid& g, std::string& err);
bool read_vrsg(const std::string& path, HostGrid& out, std::string& err);
bool make_procedural(int kind, uint32_t resolution, HostGrid& out, std::string& err);

}  // namespace vrs
