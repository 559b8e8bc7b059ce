#include "bb_engine.cuh"
namespace bb { EngineBase *make_engine_f32(const bb_desc &d) { return new Engine<float>(d); } }
