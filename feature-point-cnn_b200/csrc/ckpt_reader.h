// Reader for the reference's checkpoints without libtorch.
//
// The reference writes its snapshots with torch.save of a dict
// {'epoch', 'model_state_dict', 'optimizer_state_dict', 'scaler_state_dict'}
// (reference python/src/saveutils.py:54-63) and reads 'model_state_dict' back for inference
// (python/src/saveutils.py:6-18).  On disk that is a ZIP archive with STORED entries:
// <root>/data.pkl (pickle protocol 2, storages as persistent ids) and <root>/data/<n> raw
// little-endian storages.  This reader parses exactly that: a minimal ZIP walker plus a small
// pickle VM that understands the opcodes and the four globals torch emits for tensors.
// A bare state_dict (what InferenceWrapper.trace saves as *_params.pt,
// python/src/inferencewrapper.py:89-91) is accepted as well.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace spb200 {

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;      // contiguous, converted to fp32 (integer tensors too)
    bool is_integer = false;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        return n;
    }
};

using StateDict = std::map<std::string, HostTensor>;

// Returns true and fills `out` with every tensor of the checkpoint's model_state_dict (or of the
// top-level dict when there is no such key).  On failure returns false and sets `err`.
bool read_checkpoint(const std::string& path, StateDict& out, std::string& err);

// InferenceWrapper.trace writes *_params.pt with the first component of every key stripped
// (python/src/inferencewrapper.py:89-91: 'encoder.conv1.weight' -> 'conv1.weight', 'detector.layer.0...' ->
// 'layer.0...').  The stripped names are still unambiguous; this puts the module prefixes back when the dict has none.
void restore_module_prefixes(StateDict& sd);

}  // namespace spb200
