{
  "targets": [
    {
      "target_name": "zkattest",
      "sources": ["zkattest_napi.cc"],
      "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "../../include"],
      "dependencies": ["<!(node -p \"require('node-addon-api').gyp\")"],
      "defines": ["NAPI_CPP_EXCEPTIONS"],
      "cflags_cc": ["-std=c++17", "-fexceptions"],
      "libraries": ["-L<(module_root_dir)/../../zkp_ecdsa_b200", "-lzkattest", "-Wl,-rpath,<(module_root_dir)/../../zkp_ecdsa_b200"]
    },
    {
      "target_name": "zkattest_war256",
      "sources": ["zkattest_napi.cc"],
      "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "../../include"],
      "dependencies": ["<!(node -p \"require('node-addon-api').gyp\")"],
      "defines": ["NAPI_CPP_EXCEPTIONS", "ZKA_NAPI_MODULE=zkattest_war256"],
      "cflags_cc": ["-std=c++17", "-fexceptions"],
      "libraries": ["-L<(module_root_dir)/../../zkp_ecdsa_b200", "-lzkattest_war256", "-Wl,-rpath,<(module_root_dir)/../../zkp_ecdsa_b200"]
    }
  ]
}
