//! add(), add_op.rs:16-107 -> b200_add: rank-4 + [C,1,1] initializer (:75) or rank-2 + rank-2 (:84); input 2 must be an
//! initializer (:54-68).  The 2-D result is printed like upstream (:94-105, 1-based class).
use std::ptr;

use onnx_protobuf::{NodeProto, TensorProto, ValueInfoProto};

use crate::device::{check, default_context, DeviceTensor, Store};
use crate::inference_engine::utils::{already_into_initializer, get_stored_tensor};

pub fn add(output_container: &Store,
           node: &NodeProto,
           model_inputs: &Vec<ValueInfoProto>,
           model_initializers: &Vec<TensorProto>) {
    let x = if already_into_initializer(model_initializers, &node.input[0]) {
        get_stored_tensor(0, node, model_inputs, model_initializers)
    } else {
        let map = output_container.lock().unwrap();
        let v = map.get(&node.input[0]).unwrap();                       // add_op.rs:41
        v.0.clone().or_else(|| v.1.clone()).expect("Cannot retrieve input 1 for Add operation from hashmap input/output")
    };
    if !already_into_initializer(model_initializers, &node.input[1]) {
        panic!("Cannot retrieve input 2 for Add operation");            // add_op.rs:66
    }
    let b = get_stored_tensor(1, node, model_inputs, model_initializers);
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { b200rt_sys::b200_add(ctx.raw(), x.raw(), b.raw(), &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    let y = DeviceTensor::from_raw(ctx.clone(), y);
    if x.rank() == 4 {
        output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(y)));
    } else {
        let result = y.to_array2();
        let (mut best, mut best_v) = (0usize, f32::MIN);
        for (i, v) in result.row(0).iter().enumerate() { if *v > best_v { best = i; best_v = *v; } }
        println!("\nMNist-8 Inference results: Class {}-nth predicted.\nActual Data: {:?}", best + 1, result);
        output_container.lock().unwrap().insert(node.output[0].clone(), (Some(y), None));
    }
}
