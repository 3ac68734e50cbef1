//! global_average_pool(), global_average_pool_op.rs:11-52 -> b200_global_avgpool.
use onnx_protobuf::NodeProto;

use super::{slot4, unary};
use crate::device::Store;

pub fn global_average_pool(output_container: &Store, node: &NodeProto) {
    let x = slot4(output_container, &node.input[0], "GlobalAveragePool");
    let y = unary(b200rt_sys::b200_global_avgpool, &x);
    output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(y)));
}
