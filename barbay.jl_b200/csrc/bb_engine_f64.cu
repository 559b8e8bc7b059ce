#include "bb_multi.cuh"
namespace bb {
EngineBase *make_engine_f64(const bb_desc &d) { return new Engine<double>(d); }
EngineBase *make_multi_engine_f64(const bb_desc &d) { return new MultiEngine<double>(d); }
}  // namespace bb
