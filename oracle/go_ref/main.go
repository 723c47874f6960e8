// go_ref - regenerates the golden fixtures of tests/golden/go/ from the REAL reference path.
//
// Test infrastructure (oracle/), not product.  There is no Go toolchain in the build image, so this program
// has never been run there: it exists so that the first box with Go turns the oracle's "restated from upstream
// behaviour" into "pinned byte for byte".  It makes exactly the library calls the reference makes on the
// tools/call path, in the same order, on the cases of tests/golden/go/cases.jsonl:
//
//   request side   json.Unmarshal into mcp.JSONRPCRequest            (pkg/server/handler.go:83-88)
//                  mcp.NewValidator().ValidateRequest / ValidateToolCallParams (handler.go:91-95, 217-219)
//                  json.Marshal(params["arguments"])                  (handler.go:224-231)
//                  dynamicpb.NewMessage + protojson.Unmarshal         (pkg/grpc/reflection.go:351-357)
//                  proto.Marshal                                      (what conn.Invoke does, reflection.go:373)
//   reply side     proto.Unmarshal into dynamicpb                     (reflection.go:363,373)
//                  protojson.Marshal                                  (reflection.go:381)
//                  json.NewEncoder(w).Encode(JSONRPCResponse{Result: ToolCallResult{TextContent(text)}})
//                                                                     (handler.go:265-270, 290-297)
//
// Usage:  go run . -fds ../../tests/golden/schemas.binpb -in ../../tests/golden/go/cases.jsonl \
//                  -out ../../tests/golden/go/golden.jsonl
//
// Two outputs of Go are not deterministic and are recorded as such: proto.Marshal of a dynamicpb message walks a
// Go map unless Deterministic is set (the golden wire is the Deterministic one = order.LegacyFieldOrder, the
// engine's GGR_ORDER_GO_LEGACY), and protojson adds a space after commas depending on a per-binary hash
// ("comma_space" in the header line says which variant this binary produces).
package main

import (
	"bufio"
	"bytes"
	"encoding/hex"
	"encoding/json"
	"flag"
	"fmt"
	"os"
	"strings"

	"github.com/aalobaidi/ggRMCP/pkg/mcp"
	"google.golang.org/protobuf/encoding/protojson"
	"google.golang.org/protobuf/proto"
	"google.golang.org/protobuf/reflect/protodesc"
	"google.golang.org/protobuf/reflect/protoreflect"
	"google.golang.org/protobuf/reflect/protoregistry"
	"google.golang.org/protobuf/types/descriptorpb"
	"google.golang.org/protobuf/types/dynamicpb"
)

// one line of cases.jsonl
type testCase struct {
	ID      string `json:"id"`
	Kind    string `json:"kind"`              // "args" | "reply" | "body"
	Message string `json:"message,omitempty"` // full name of the message (args, reply)
	Args    string `json:"args,omitempty"`    // kind args: the JSON a client put into params.arguments (base64 when args_b64)
	ArgsB64 bool   `json:"args_b64,omitempty"`
	WireHex string `json:"wire_hex,omitempty"` // kind reply: what the backend sent
	ReqID   string `json:"req_id,omitempty"`   // kind reply: JSON text of the request id
	Body    string `json:"body,omitempty"`     // kind body: a whole HTTP request body (base64)
}

// one line of golden.jsonl
type golden struct {
	ID        string `json:"id"`
	Error     string `json:"error,omitempty"`      // where the reference stops, with Go's wording
	Stage     string `json:"stage,omitempty"`      // decode | validate | params | marshal_args | protojson | proto_unmarshal | ok
	CanonArgs string `json:"canon_args,omitempty"` // base64 of json.Marshal(arguments)
	WireHex   string `json:"wire_hex,omitempty"`   // proto.MarshalOptions{Deterministic: true}
	JSON      string `json:"json,omitempty"`       // base64 of protojson.Marshal
	HTTPBody  string `json:"http_body,omitempty"`  // base64 of the Encode()d response
	Tool      string `json:"tool,omitempty"`
}

func b64(b []byte) string { return encodeB64(b) }

func main() {
	fdsPath := flag.String("fds", "../../tests/golden/schemas.binpb", "FileDescriptorSet")
	inPath := flag.String("in", "../../tests/golden/go/cases.jsonl", "cases")
	outPath := flag.String("out", "../../tests/golden/go/golden.jsonl", "golden output")
	flag.Parse()

	raw, err := os.ReadFile(*fdsPath)
	must(err)
	var fds descriptorpb.FileDescriptorSet
	must(proto.Unmarshal(raw, &fds))
	files, err := protodesc.NewFiles(&fds)
	must(err)
	findMsg := func(name string) (protoreflect.MessageDescriptor, error) {
		d, err := files.FindDescriptorByName(protoreflect.FullName(name))
		if err != nil {
			return nil, err
		}
		md, ok := d.(protoreflect.MessageDescriptor)
		if !ok {
			return nil, fmt.Errorf("%s is not a message", name)
		}
		return md, nil
	}
	// tool name -> input message, as pkg/types/service.go:53-61 builds the names (reflection route: full package)
	tools := map[string]protoreflect.MessageDescriptor{}
	files.RangeFiles(func(fd protoreflect.FileDescriptor) bool {
		for i := 0; i < fd.Services().Len(); i++ {
			sd := fd.Services().Get(i)
			for j := 0; j < sd.Methods().Len(); j++ {
				m := sd.Methods().Get(j)
				name := strings.ToLower(strings.ReplaceAll(string(sd.FullName()), ".", "_")) + "_" + strings.ToLower(string(m.Name()))
				tools[name] = m.Input()
			}
		}
		return true
	})
	_ = protoregistry.GlobalTypes

	in, err := os.Open(*inPath)
	must(err)
	defer in.Close()
	out, err := os.Create(*outPath)
	must(err)
	defer out.Close()
	w := bufio.NewWriter(out)
	defer w.Flush()
	enc := json.NewEncoder(w)
	enc.SetEscapeHTML(false)

	// header line: which protojson variant this binary produces
	probe, _ := findMsg("hello.HelloRequest")
	pm := dynamicpb.NewMessage(probe)
	must(protojson.Unmarshal([]byte(`{"name":"a","email":"b"}`), pm))
	pj, _ := protojson.Marshal(pm)
	must(enc.Encode(map[string]interface{}{"header": true, "comma_space": bytes.Contains(pj, []byte(`, "`)), "protobuf": "v1.36.6", "grpc": "v1.74.2"}))

	validator := mcp.NewValidator()
	sc := bufio.NewScanner(in)
	sc.Buffer(make([]byte, 1<<20), 64<<20)
	for sc.Scan() {
		var c testCase
		if err := json.Unmarshal(sc.Bytes(), &c); err != nil {
			continue
		}
		g := golden{ID: c.ID}
		switch c.Kind {
		case "args":
			md, err := findMsg(c.Message)
			must(err)
			argText := []byte(c.Args)
			if c.ArgsB64 {
				argText = decodeB64(c.Args)
			}
			// the client's arguments arrive as part of params: decode as the envelope decoder does, re-marshal as handleToolsCall does
			var args interface{}
			if err := json.Unmarshal(argText, &args); err != nil {
				g.Stage, g.Error = "decode", err.Error()
				break
			}
			requestSide(&g, md, args)
		case "reply":
			md, err := findMsg(c.Message)
			must(err)
			wire, _ := hex.DecodeString(c.WireHex)
			msg := dynamicpb.NewMessage(md)
			if err := proto.Unmarshal(wire, msg); err != nil {
				g.Stage, g.Error = "proto_unmarshal", err.Error()
				break
			}
			text, err := protojson.Marshal(msg) // reflection.go:381
			if err != nil {
				g.Stage, g.Error = "protojson", err.Error()
				break
			}
			g.JSON = b64(text)
			var id mcp.RequestID
			if c.ReqID != "" {
				if err := id.UnmarshalJSON([]byte(c.ReqID)); err != nil {
					g.Stage, g.Error = "decode", err.Error()
					break
				}
			}
			resp := mcp.JSONRPCResponse{JSONRPC: "2.0", ID: id, Result: &mcp.ToolCallResult{Content: []mcp.ContentBlock{mcp.TextContent(string(text))}, IsError: false}}
			var hb bytes.Buffer
			must(json.NewEncoder(&hb).Encode(resp)) // handler.go:290-297
			g.HTTPBody = b64(hb.Bytes())
			g.Stage = "ok"
		case "body":
			body := decodeB64(c.Body)
			var req mcp.JSONRPCRequest
			if err := json.Unmarshal(body, &req); err != nil { // handler.go:83-88
				g.Stage, g.Error = "decode", err.Error()
				break
			}
			if err := validator.ValidateRequest(&req); err != nil { // handler.go:91-95
				g.Stage, g.Error = "validate", err.Error()
				break
			}
			if req.Method != "tools/call" {
				g.Stage, g.Error = "validate", "method "+req.Method
				break
			}
			if err := validator.ValidateToolCallParams(req.Params); err != nil { // handler.go:217-219
				g.Stage, g.Error = "params", err.Error()
				break
			}
			tool, _ := req.Params["name"].(string)
			g.Tool = tool
			md, ok := tools[tool]
			if !ok {
				g.Stage, g.Error = "params", "tool "+tool+" not found"
				break
			}
			requestSide(&g, md, req.Params["arguments"])
		}
		must(enc.Encode(&g))
	}
}

// handler.go:224-231 + reflection.go:351-357 + proto.Marshal
func requestSide(g *golden, md protoreflect.MessageDescriptor, args interface{}) {
	var argumentsJSON string
	if args != nil {
		b, err := json.Marshal(args)
		if err != nil {
			g.Stage, g.Error = "marshal_args", err.Error()
			return
		}
		argumentsJSON = string(b)
	}
	g.CanonArgs = b64([]byte(argumentsJSON))
	msg := dynamicpb.NewMessage(md)
	if argumentsJSON != "" && argumentsJSON != "{}" { // reflection.go:354
		if err := protojson.Unmarshal([]byte(argumentsJSON), msg); err != nil {
			g.Stage, g.Error = "protojson", err.Error()
			return
		}
	}
	wire, err := proto.MarshalOptions{Deterministic: true}.Marshal(msg)
	if err != nil {
		g.Stage, g.Error = "proto_marshal", err.Error()
		return
	}
	g.WireHex = hex.EncodeToString(wire)
	g.Stage = "ok"
}

func must(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, "go_ref:", err)
		os.Exit(1)
	}
}
