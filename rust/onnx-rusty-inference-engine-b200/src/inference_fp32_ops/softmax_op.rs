//! softmax(), softmax_op.rs:13-57 -> b200_softmax.  Like upstream (:30-41) the result is printed (1-based class) and NOT
//! inserted in the store.
use onnx_protobuf::NodeProto;

use super::{slot4, unary};
use crate::device::Store;

pub fn softmax(output_container: &Store, node: &NodeProto) {
    let x = slot4(output_container, &node.input[0], "Softmax");
    let result = unary(b200rt_sys::b200_softmax, &x).to_array2();
    let row = result.row(0);
    let (mut best, mut best_v) = (0usize, f32::MIN);
    for (i, v) in row.iter().enumerate() { if *v > best_v { best = i; best_v = *v; } }
    println!("\nSqueezenet1.0-8 Inference results: Class {}-nth predicted.\nActual Data: {:?}", best + 1, result);
}
