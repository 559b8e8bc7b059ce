#include "bb_engine.cuh"
namespace bb { EngineBase *make_engine_f64(const bb_desc &d) { return new Engine<double>(d); } }
