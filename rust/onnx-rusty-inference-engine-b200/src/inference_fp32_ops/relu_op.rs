//! relu(), relu_op.rs:11-33 -> b200_relu.
use onnx_protobuf::NodeProto;

use super::{slot4, unary};
use crate::device::Store;

pub fn relu(output_container: &Store, node: &NodeProto) {
    let x = slot4(output_container, &node.input[0], "Relu");
    let y = unary(b200rt_sys::b200_relu, &x);
    output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(y)));
}
