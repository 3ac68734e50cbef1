//! inference() / node_inference(), model_inference.rs:29-162, over the B200 backend.
use std::collections::HashMap;
use std::sync::{Arc, Mutex};

use onnx_protobuf::{ModelProto, NodeProto};
use protobuf::Message;

use crate::device::{default_context, Model, Store};
use crate::inference_engine::utils::manage_input_data;
use crate::inference_fp32_ops::add_op::add;
use crate::inference_fp32_ops::concatenate_op::concatenation;
use crate::inference_fp32_ops::convolution_op::convolution;
use crate::inference_fp32_ops::dropout_op::drop_out;
use crate::inference_fp32_ops::global_average_pool_op::global_average_pool;
use crate::inference_fp32_ops::max_pool_op::max_pool;
use crate::inference_fp32_ops::mul_op::mul;
use crate::inference_fp32_ops::relu_op::relu;
use crate::inference_fp32_ops::reshape_op::reshape;
use crate::inference_fp32_ops::softmax_op::softmax;

/// model_inference.rs:29.  Same signature.  The node walk is in file order on one CUDA stream: upstream's branch threads
/// (multithreading/*.rs) change which host thread issues a node, never a result.  Activations stay in HBM in the store.
pub fn inference(model: ModelProto, input_data: Vec<f32>, input_tensor_name: Vec<&str>) {
    let hashmap_outputs_to_inputs: Store = Arc::new(Mutex::new(HashMap::new()));
    let arc_model = Arc::new(model);
    manage_input_data(&hashmap_outputs_to_inputs, &arc_model.graph.input, &arc_model.graph.initializer, &input_data, &input_tensor_name);
    for node in &arc_model.graph.node {
        node_inference(node, &hashmap_outputs_to_inputs, &arc_model);
    }
}

/// The graph-level path (what the benchmarks time): the whole walk planned once, fused, replayed as a CUDA graph.
/// Returns the output rows instead of printing them; `batch` images in `input_data`.
pub fn inference_fused(model: &ModelProto, input_data: &[f32], batch: i64) -> Vec<f32> {
    let bytes = model.write_to_bytes().expect("serialise ModelProto");
    let m = Model::from_bytes(default_context(), &bytes).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    m.run(input_data, batch).unwrap_or_else(|e| panic!("b200rt: {}", e.message))
}

/// model_inference.rs:128-162: the op-dispatch boundary, same ten names, same panic on anything else (:158).
pub fn node_inference(node: &NodeProto, hashmap_outputs_to_inputs: &Store, model: &Arc<ModelProto>) {
    println!("INFERENCE ON INPUT(s) {:?} ON {} OPERATION done by {}", node.input, node.op_type.clone(),
             std::thread::current().name().unwrap_or("MAIN PROCESS"));
    let operation = &node.op_type;
    match operation.as_str() {
        "Conv" => convolution(hashmap_outputs_to_inputs, node, &model.graph.input, &model.graph.initializer),
        "Relu" => relu(hashmap_outputs_to_inputs, node),
        "MaxPool" => max_pool(hashmap_outputs_to_inputs, node, &model.graph.input, &model.graph.initializer),
        "Concat" => concatenation(hashmap_outputs_to_inputs, node),
        "Dropout" => drop_out(hashmap_outputs_to_inputs, node),
        "GlobalAveragePool" => global_average_pool(hashmap_outputs_to_inputs, node),
        "Softmax" => softmax(hashmap_outputs_to_inputs, node),
        "Reshape" => reshape(hashmap_outputs_to_inputs, node, &model.graph.input, &model.graph.initializer),
        "Add" => add(hashmap_outputs_to_inputs, node, &model.graph.input, &model.graph.initializer),
        "MatMul" => mul(hashmap_outputs_to_inputs, node),
        _ => panic!("INFERENCE OPERATION '{}' NOT FOUND FOR NODE {}", operation.as_str(), <String as AsRef<str>>::as_ref(&node.name)),
    }
}
