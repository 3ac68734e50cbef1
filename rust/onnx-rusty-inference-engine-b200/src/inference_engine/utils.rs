//! utils.rs:14-197 over device tensors.
use onnx_protobuf::type_proto::Value;
use onnx_protobuf::tensor_shape_proto::dimension::Value::{DimParam, DimValue};
use onnx_protobuf::{NodeProto, TensorProto, ValueInfoProto};

use crate::device::{cached_initializer, default_context, DeviceTensor, Store};

/// utils.rs:14-21
pub fn already_into_initializer(model_initializers: &Vec<TensorProto>, input_name: &str) -> bool {
    model_initializers.iter().any(|t| <String as AsRef<str>>::as_ref(&t.name) == input_name)
}

/// utils.rs:192-197
pub fn u8_to_f32(bytes: &[u8]) -> f32 { f32::from_le_bytes([bytes[0], bytes[1], bytes[2], bytes[3]]) }

/// utils.rs:53-97: static dims of a graph.input entry; DimParam panics like upstream (:67).
pub fn get_input_data_shape(model_inputs: &Vec<ValueInfoProto>, input_name: &str) -> Vec<i64> {
    let mut shape = vec![];
    for inp in model_inputs {
        if <String as AsRef<str>>::as_ref(&inp.name) != input_name { continue; }
        match inp.type_.value.as_ref().expect("UNABLE TO RETRIEVE TYPE OF NODE") {
            Value::TensorType(t) => for el in &t.shape.as_ref().expect("UNABLE TO RETRIEVE SHAPE VALUES").dim {
                match el.value.as_ref().expect("UNABLE TO RETRIEVE DIMS VALUES") {
                    DimValue(v) => shape.push(*v),
                    DimParam(_) => panic!("DIM PARAM NOT YET IMPLEMENTED FOR NODE {}'S INPUT", input_name),
                    _ => panic!("SHAPE DIMS UNRECOGNIZED FOR NODE {}'S INPUT", input_name),
                }
            },
            _ => panic!("UNRECOGNIZED TYPE FOR NODE {}'S INPUT", input_name),
        }
        break;
    }
    shape
}

fn initializer<'a>(name: &str, model_initializers: &'a Vec<TensorProto>) -> &'a TensorProto {
    model_initializers.iter().find(|t| <String as AsRef<str>>::as_ref(&t.name) == name).unwrap_or_else(|| panic!("initializer {name} not found"))
}

/// The int64 payload of an initializer (Reshape's shape, utils.rs:138-142 / :170-183).
pub fn initializer_i64(name: &str, model_initializers: &Vec<TensorProto>) -> Vec<i64> {
    let t = initializer(name, model_initializers);
    if !t.raw_data.is_empty() { t.raw_data.chunks_exact(8).map(|c| i64::from_le_bytes(c.try_into().unwrap())).collect() } else { t.int64_data.clone() }
}

/// get_stored_tensor (utils.rs:113-185) for f32 initializers: raw_data as LE f32 chunks (:128-133) else float_data
/// (:134-137); rank from graph.input's dims (:122), else TensorProto.dims.  Uploaded once, not once per use.
pub fn get_stored_tensor(input_index: usize, node: &NodeProto, model_inputs: &Vec<ValueInfoProto>, model_initializers: &Vec<TensorProto>) -> DeviceTensor {
    let name = &node.input[input_index];
    let t = initializer(name, model_initializers);
    let mut dims = get_input_data_shape(model_inputs, name);
    if dims.is_empty() { dims = t.dims.clone(); }
    if dims.is_empty() || dims.len() > 4 { panic!("initializer {name}: rank {} unsupported (utils.rs:146-184)", dims.len()); }
    cached_initializer(t, &dims, || {
        if !t.raw_data.is_empty() { t.raw_data.chunks_exact(4).map(u8_to_f32).collect() } else { t.float_data.clone() }
    })
}

/// manage_input_data, utils.rs:29-45: names that are initializers are skipped (:35); every other name receives the same
/// input_data shaped by graph.input's static dims (:36-40; a length mismatch panics like from_shape_vec().unwrap()).
pub fn manage_input_data(hashmap_outputs_to_inputs: &Store, model_inputs: &Vec<ValueInfoProto>, model_initializers: &Vec<TensorProto>,
                         input_data: &Vec<f32>, input_tensor_name: &Vec<&str>) {
    for name in input_tensor_name {
        if already_into_initializer(model_initializers, name) { continue; }
        let dims = get_input_data_shape(model_inputs, name);
        assert!(dims.len() == 4 && dims.iter().product::<i64>() as usize == input_data.len(), "input length != static model shape (utils.rs:40)");
        let t = default_context().upload(&dims, input_data).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
        hashmap_outputs_to_inputs.lock().unwrap().insert(name.to_string(), (None, Some(t)));
    }
}
