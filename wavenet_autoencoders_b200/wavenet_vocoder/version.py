version = "0.2.0+b200"
