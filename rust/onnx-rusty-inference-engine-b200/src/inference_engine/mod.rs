pub mod model_inference;
pub mod utils;
// The reference's `multithreading` module (branch threads per Fire module) has no counterpart: node results are
// order-independent, all work is stream-ordered on the device, and the graph-level path replays one CUDA graph.
