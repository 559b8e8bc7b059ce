#include "bb_multi.cuh"
namespace bb {
EngineBase *make_engine_f32(const bb_desc &d) { return new Engine<float>(d); }
EngineBase *make_multi_engine_f32(const bb_desc &d) { return new MultiEngine<float>(d); }
}  // namespace bb
