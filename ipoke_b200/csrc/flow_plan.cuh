// Logical op list of the conditional MaCow flow (shared by the inference plan, flow.cu, and the training plan, flow_train.cu):
// exact mirror of the reference's module order, MultiScaleInternal.forward (macow2.py:873-920) / MaCowStep (macow2.py:1066-1117) /
// MaCowUnit (macow2.py:957-995) / MultiScalePrior (macow2.py:569-593).
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace ipk {

enum LKind { L_ACTNORM, L_SHUFFLE, L_MCF, L_NICE };
struct LogicalOp {
  LKind kind;
  int C;               // active channels of the level
  std::string prefix;  // state-dict prefix
  int order = 0;       // MCF order
  int coff = 0, cnt = 0;  // actnorm range
  bool fwd_idx = false;   // shuffle: use forward_shuffle_idx
  int factor = 2; bool skip = false; bool up = true;  // NICE
  int nice_id = -1;
};


struct LevelInfo { int L, C, steps, prior_factor, prior_out, z1; };

// MultiScaleInternal.__init__ channel bookkeeping (macow2.py:825-871)
std::vector<LevelInfo> levels_of(const ipk_flow_config& c);
// channel index lists of NICE2d.split (macow2.py:301-317,364-377): iz = network input, ip = transformed part
void nice_indices(int C, int factor, bool skip, bool up, std::vector<int>& iz, std::vector<int>& ip);
// ops in execution order of the forward (density) direction, or of the inverse
std::vector<LogicalOp> logical_program(const ipk_flow_config& cfg, bool fwd);

}  // namespace ipk
